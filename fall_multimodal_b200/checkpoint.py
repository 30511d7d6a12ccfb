"""Reference checkpoint formats (the part adjacent to the modules).

``F2/main.py:318-341`` writes ``best_model.pt`` = ``{"model_weight": state_dict}`` and ``checkpoint.pt`` =
``{"epoch", "model_weight", "optimizer", "lr_scheduler", "scaler", "best_acc"}``; the notebooks save bare
state_dicts (``tsstg-model_best.pth``) of classes whose attribute names differ from the package modules
(``pts_stream`` / ``mot_stream`` / ``sensor`` / ``fcn``, ``st_gcn_networks``). Everything here is key plumbing on
the host: the tensors are the reference's own, shapes are checked by ``load_state_dict``.
"""
from __future__ import annotations

import torch

# (reference name, name in this package); applied to whole dotted components
_RENAMES = (
    ("pts_stream", "stgcan_1"),
    ("mot_stream", "stgcan_2"),
    ("fcn", "fc"),
    ("st_gcn_networks", "st_gcan_networks"),
)


def _rename(key: str, sensor_attr: str | None) -> str:
    parts = key.split(".")
    if parts and parts[0] == "module":          # nn.DataParallel / DistributedDataParallel wrappers
        parts = parts[1:]
    ren = dict(_RENAMES)
    out = [ren.get(p, p) for p in parts]
    if sensor_attr and out and out[0] == "sensor":
        out[0] = sensor_attr
    return ".".join(out)


def remap_reference_state_dict(sd: dict, model: torch.nn.Module) -> dict:
    """Rename the keys of a reference / notebook state_dict to the names ``model`` uses.

    Keys that already match are left alone; the notebook's ``sensor.`` branch maps to whichever of ``lstm`` / ``cnn``
    the target model has. Unknown keys are passed through so that ``load_state_dict(strict=True)`` reports them."""
    want = set(model.state_dict().keys())
    sensor_attr = next((a for a in ("lstm", "cnn", "sensor") if any(k.startswith(a + ".") for k in want)), None)
    out = {}
    for k, v in sd.items():
        out[k if k in want else _rename(k, sensor_attr)] = v
    return out


def load_reference_checkpoint(model: torch.nn.Module, path_or_obj, strict: bool = True, map_location="cpu",
                              allow_pickle: bool = False) -> dict:
    """Load ``best_model.pt`` / ``checkpoint.pt`` / a bare notebook state_dict into ``model``.

    ``path_or_obj``: a path (``str`` / ``bytes`` / ``os.PathLike``), an open file, or an already loaded object. Files are read
    with ``weights_only=True`` (tensors and plain containers only); a checkpoint that needs arbitrary pickled classes (e.g. a
    pickled lr-scheduler object) is refused unless ``allow_pickle=True`` - unpickling runs code from the file.

    Returns the rest of the checkpoint (``epoch``, ``optimizer``, ``best_acc`` ... when present) so a resume can
    restore the optimizer exactly as ``F2/main.py:294-303`` does."""
    import os
    import pickle

    if isinstance(path_or_obj, (str, bytes, os.PathLike)) or hasattr(path_or_obj, "read"):
        try:
            obj = torch.load(path_or_obj, map_location=map_location, weights_only=True)
        except pickle.UnpicklingError as e:
            if not allow_pickle:
                raise RuntimeError(f"checkpoint needs full unpickling ({e}); pass allow_pickle=True if you trust the file") from e
            if hasattr(path_or_obj, "seek"):
                path_or_obj.seek(0)
            obj = torch.load(path_or_obj, map_location=map_location, weights_only=False)
    else:
        obj = path_or_obj
    sd = obj["model_weight"] if isinstance(obj, dict) and "model_weight" in obj else obj
    model.load_state_dict(remap_reference_state_dict(sd, model), strict=strict)
    return {k: v for k, v in obj.items() if k != "model_weight"} if isinstance(obj, dict) and "model_weight" in obj else {}


def save_checkpoint(path, model: torch.nn.Module, optimizer=None, lr_scheduler=None, scaler=None, epoch: int | None = None,
                    best_acc: float | None = None) -> None:
    """Write the reference's layout: ``{"model_weight": ...}`` alone (best_model.pt) or with the resume fields."""
    ckpt = {"model_weight": model.state_dict()}
    if optimizer is not None or epoch is not None:
        ckpt.update({"epoch": epoch, "optimizer": optimizer.state_dict() if optimizer is not None else None,
                     "lr_scheduler": lr_scheduler.state_dict() if lr_scheduler is not None else None,
                     "scaler": scaler.state_dict() if scaler is not None else None, "best_acc": best_acc})
    torch.save(ckpt, path)

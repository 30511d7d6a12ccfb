"""Golden vectors for the import-time EmbGCN variants (EmbGCN.py:91-123: EmbGCN_noGate, EmbGCN_linear) from the UNMODIFIED
reference classes: inputs, parameters and outputs of one forward each -> tests/golden/embgcn_variants.pt.
Run here (needs /root/reference): ``python -m oracle.make_golden_variants``."""
import os
import sys

import torch

from oracle import ref_import


def main():
    ref_import.load_tragcn(30)
    mod = sys.modules["TRAGCN.EmbGCN"]
    g = torch.Generator().manual_seed(11)
    V, Din, Dout, Ed, B = 14, 10, 16, 8, 3
    adj = torch.ones(V, V)
    out = {}
    for name in ("noGate", "linear"):
        m = getattr(mod, "EmbGCN_" + name)(Din, Dout, adj, 2, Ed)
        sd = {k: torch.randn(v.shape, generator=g) * 0.3 for k, v in m.state_dict().items()}
        m.load_state_dict(sd)
        x = torch.randn(B, V, Din, generator=g)
        E = torch.randn(V, Ed, generator=g)
        with torch.no_grad():
            y = m(x, E)
        out[name] = {"state_dict": sd, "x": x, "E": E, "y": y, "keys": list(sd.keys())}
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "embgcn_variants.pt")
    torch.save(out, dst)
    print("wrote", dst, {k: tuple(v["y"].shape) for k, v in out.items()})


if __name__ == "__main__":
    main()

"""How far does a free-running bf16 decode stay within the 1e-2 tolerance of the fp64 oracle?  Free running feeds every rounding
difference back, so the comparison prefix N of the parity tests is a choice; this prints the running maximum error over time for the
batch-256 kernel and the small-batch kernel on the same inputs (external randomness).   python tools/drift_vs_oracle.py [B] [Tv] [T]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gst_tacotron_b200.runtime import Engine  # noqa: E402
from oracle import reference_port as O  # noqa: E402
from tests.util import make_cfg, make_weights, oracle_decode  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    Tv = int(sys.argv[2]) if len(sys.argv) > 2 else 82
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    cfg = make_cfg("SMA", precision="bf16")
    W = make_weights(cfg)
    eng = Engine(cfg, W)
    enc, _, k0, k1, nz = O.synth_decoder_inputs(cfg, B, Tv, T, teacher=False)
    ref = oracle_decode(cfg, W, enc, steps=T, keep0=k0, keep1=k1, noise=nz)
    marks = [t for t in (8, 16, 32, 64, 100, 150, 200, 300, 400, 600, 1000) if t <= T]
    for kernel in ("batch", "small"):
        out = eng.decode(encodings=enc, steps=T, rng="external", keep0=k0, keep1=k1, noise=nz, kernel=kernel)
        em = np.abs(out["mel"] - ref["decodings"]).max(axis=(0, 2))
        es = np.abs(out["stop"] - ref["stops"]).max(axis=0)
        ea = np.abs(out["alignment"] - ref["alignments"]).max(axis=(0, 2))
        flips = ((out["stop"] < 0) != (ref["stops"] < 0)) & (np.abs(ref["stops"]) > 1e-2)
        print("kernel %-5s  B=%d Tv=%d: running max |error| up to step t (mel / stop / alignment), decidable stop-sign flips" % (kernel, B, Tv))
        for t in marks:
            print("   t <= %4d: %.2e / %.2e / %.2e   flips %d" % (t, em[:t].max(), es[:t].max(), ea[:t].max(), int(flips[:, :t].sum())))
    eng.close()


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY: (re)generate tests/golden/*.npz.

Run in the build container (needs /root/reference and `make -C oracle ref`):
    python oracle/make_golden.py
Two kinds of fixtures:
  cavity_sample.npz  -- the reference's OWN golden sample (examples/cavity/sample/
                        exp.iter_history.plt and exp.field.0.vts), stored as the printed
                        6-significant-digit tokens so tests can compare text-exactly.
  ref_<case>.npz     -- raw fp64 fields, SIMPLE residual history and linear-solver sweep
                        counts dumped from the real reference (oracle/_ref/ref_dump, built
                        with -ffp-contract=off) for the small cases in tests/cases.py.
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cases  # noqa: E402
import refrun  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SAMPLE = "/root/reference/examples/cavity/sample"

REF_CASES = cases.GOLDEN_CASES


def cavity_sample():
    txt = open(os.path.join(SAMPLE, "exp.iter_history.plt")).read().splitlines()
    rows = [l.split() for l in txt if re.match(r"^\d+ ", l)]
    rs_tokens = np.array([r[1] for r in rows])
    vts = open(os.path.join(SAMPLE, "exp.field.0.vts")).read()
    arrays = {}
    for m in re.finditer(r'<DataArray Name="(\w+)"[^>]*>\n(.*?)\n\s*</DataArray>', vts, re.S):
        arrays[m.group(1)] = np.array(m.group(2).split())
    np.savez_compressed(os.path.join(GOLD, "cavity_sample.npz"), rs_tokens=rs_tokens, vts_text=np.array(vts),
                        velocity_x=arrays["velocity_x"], velocity_y=arrays["velocity_y"],
                        pressure=arrays["pressure"], volume_fraction_0=arrays["volume_fraction_0"])
    print("cavity_sample: %d residuals, %d cells" % (len(rs_tokens), len(arrays["pressure"])))


def main():
    os.makedirs(GOLD, exist_ok=True)
    cavity_sample()
    if not refrun.available():
        raise SystemExit("oracle/_ref/ref_dump missing: run `make -C oracle ref` first")
    only = os.environ.get("GOLDEN_ONLY")   # comma-separated case names: (re)generate just these
    for name, (p, nsteps) in REF_CASES.items():
        if only and name not in only.split(","):
            continue
        res, _ = refrun.run_ref_dump(p, nsteps, iters=True)
        # the reference's own step() must give the same fields as the --iters call sequence
        res2, _ = refrun.run_ref_dump(p, nsteps, iters=False)
        for k in ("u0", "p", "flux", "pd0"):
            assert np.array_equal(res[k], res2[k]), (name, k)
        keep = {k: v for k, v in res.items() if not k.startswith("stforce")}
        np.savez_compressed(os.path.join(GOLD, "ref_%s.npz" % name), nsteps=nsteps, **keep)
        print("ref_%s: %d cells, %d SIMPLE its, %d linear solves" %
              (name, len(res["p"]), len(res["rs"]), len(res["lin_iters"])))


if __name__ == "__main__":
    main()

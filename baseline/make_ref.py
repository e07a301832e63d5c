#!/usr/bin/env python
"""Recipe for the reference arm of bench.py.

Copies the reference's two sampler modules, UNMODIFIED, from /root/reference into the git-ignored `baseline/_ref/`
(it is not product source and never enters the history; it travels to the GPU box with the working tree, where
/root/reference does not exist).  `bench.py --impl reference` imports them from there and times the reference's own
`mcmc_draw_parameters` on the box's host cores.  Nothing under mcmc_clv_model_b200/, src/ or tests/ touches them.

    python baseline/make_ref.py            # no-op (exit 0) when /root/reference is absent
"""
import hashlib
import os
import shutil
import sys

REF = os.environ.get("CLV_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = {"src/models/bivariate/mcmc.py": "bivariate_mcmc.py", "src/models/trivariate/mcmc.py": "trivariate_mcmc.py"}


def main():
    if not os.path.isdir(REF):
        print(f"{REF} not present: baseline/_ref left as it is")
        return 0
    os.makedirs(OUT, exist_ok=True)
    lines = []
    for src, dst in FILES.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(OUT, dst))
        lines.append(f"{hashlib.sha256(open(os.path.join(OUT, dst), 'rb').read()).hexdigest()}  {dst}  <- {src}")
    with open(os.path.join(OUT, "MANIFEST.txt"), "w") as f:
        f.write("unmodified copies made by baseline/make_ref.py\n" + "\n".join(lines) + "\n")
    print("\n".join(lines))
    return 0


if __name__ == "__main__":
    sys.exit(main())

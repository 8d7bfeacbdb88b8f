#!/usr/bin/env python
"""Stage the reference files the on-box comparisons need under baseline/_ref/ (git-ignored, NOT gpurun-ignored, so the
copy travels to the GPU box where /root/reference does not exist).  Nothing under baseline/_ref/ is product source and
nothing in lc_b200/ imports it; it is used by
  * tests/test_compat_gpu.py       the reference's unmodified lib/pnp/cer_solver.py + pnp_ceres.py on liblc_b200.so
  * tools/train_step.py            ptnet + Loss_fn.forward with the reference Loss_cov_mixed vs lc_b200 swapped in
  * tools/bench_reference_gpu.py   the reference's PyTorch-op graph (Loss_cov_mixed fwd+bwd) timed on the same B200

    python tools/stage_reference.py [--src /root/reference]
"""
import argparse
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["losses.py", "floatbits.py", "symmetry.py", "ptnet.py", "model_transform.py",
         "lib/cov_mixed.py", "lib/nll/pnp_auto.py", "lib/nll/pnp_utils.py",
         "lib/transforms/__init__.py", "lib/transforms/transforms.py", "lib/transforms/rotation_conversions.py",
         "lib/pnp/cer_solver.py", "lib/pnp/pnp_ceres.py", "lib/pnp/cv2_solver.py",
         "lib/utils/grad.py", "lib/optim/ranger.py", "lib/optim/lr_scheduler.py",
         "model/__init__.py", "model/cdpn_resnet.py", "model/zebra_DeepLabV3.py", "model/zebra_resnet.py",
         "configs/glmo.yaml", "configs/gsplmo.yaml", "configs/gycbv.yaml", "configs/zlmo.yaml", "configs/zycbv.yaml"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    dst = os.path.join(ROOT, "baseline", "_ref")
    n = 0
    for f in FILES:
        s = os.path.join(a.src, f)
        if not os.path.exists(s):
            print("missing in the reference:", f)
            continue
        d = os.path.join(dst, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        n += 1
    print(f"staged {n} files under {dst}")


if __name__ == "__main__":
    main()

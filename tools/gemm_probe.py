"""Developer tool: GEMM-stage time of the local streaming top-k (dvl_shard_topk) on a slice of BASELINE
config 4 (rows x 768 bf16, 10 k queries), for A/B runs of kernel variants selected by environment variables
(DVL_GEMM_PAIR, DVL_GEMM_PROBE = no epilogue: the MMA/TMA pipeline alone).  No result checks."""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from debias_vision_lang_b200 import _capi, synthetic  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_500_000
    q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    w = synthetic.clip_like_rows(rows, 768, q, 2, 1238, 0, rows, device="cuda", dtype=torch.bfloat16)
    lab = torch.from_numpy(w.labels.astype(np.uint8)).cuda()
    ctx = _capi.context(0)
    for _ in range(3):
        ctx.shard_topk(w.gallery, w.queries, lab, 0, 2, 1000, _capi.PRECISION_BF16, 0)
    torch.cuda.synchronize()
    ctx.get_profile()
    ctx.set_profiling(True)
    steps = 5
    for _ in range(steps):
        ctx.shard_topk(w.gallery, w.queries, lab, 0, 2, 1000, _capi.PRECISION_BF16, 0)
    torch.cuda.synchronize()
    prof = ctx.get_profile()
    gemm = prof["gemm"][0] / steps
    print({"rows": rows, "gemm_ms": round(gemm, 3), "tflops": round(2.0 * rows * q * 768 / (gemm * 1e-3) / 1e12, 1),
           "select_ms": round(prof["select"][0] / steps, 3)})


if __name__ == "__main__":
    main()

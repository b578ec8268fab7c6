"""Developer tool: per-stage device time of the full-sort ranking route on a cfg2-shaped score
matrix (random scores), without any result checks -- used for A/B runs of kernel variants,
including deliberately incomplete ones (DVL_SORT_DBG)."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from debias_vision_lang_b200 import _capi  # noqa: E402


def main():
    n, q, c = 108_501, 319, 7
    if len(sys.argv) > 2:
        n, q = int(sys.argv[1]), int(sys.argv[2])
    torch.manual_seed(0)
    scores = torch.randn(q, n, device="cuda") * 0.4 + 1.0
    labels = torch.randint(0, c, (n,), device="cuda", dtype=torch.uint8)
    labels[:c] = torch.arange(c, dtype=torch.uint8)
    ctx = _capi.context(0)
    for _ in range(5):
        ctx.rank_and_score(scores, labels, c, n)
    ctx.set_profiling(True)
    ctx.get_profile()
    steps = 20
    for _ in range(steps):
        ctx.rank_and_score(scores, labels, c, n)
    prof = ctx.get_profile()
    print({k: round(v[0] / steps, 4) for k, v in prof.items() if v[0] > 0})


if __name__ == "__main__":
    main()

import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from debias_vision_lang_b200 import _capi, measuring_bias, synthetic
w = synthetic.baseline_workload("cfg2", device="cuda")
lab_np = w.labels.astype(np.int64)
lab_dev = torch.from_numpy(lab_np).cuda()
lab_u8 = lab_dev.to(torch.uint8)
ctx = _capi.context(0)
n = w.gallery.shape[0]
def timeit(f, k=30):
    for _ in range(5): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
print("C ABI resident          ms", round(timeit(lambda: ctx.eval_ranking(w.gallery, w.queries, lab_u8, 7, n, 0)), 4))
print("evaluate(np int64 labels) ms", round(timeit(lambda: measuring_bias.evaluate(lab_np, w.gallery, w.queries, 1.0)), 4))
print("evaluate(cuda int64 labels) ms", round(timeit(lambda: measuring_bias.evaluate(lab_dev, w.gallery, w.queries, 1.0)), 4))
print("eval_ranking x2 (ref style) ms", round(timeit(lambda: (measuring_bias.eval_ranking(lab_np, w.gallery, w.queries, "maxskew"), measuring_bias.eval_ranking(lab_np, w.gallery, w.queries, "ndkl"))), 4))

"""CPU-side tests: host logic of the reference-interface mirror, and that the C-ABI library
loads and exports every symbol include/dvl_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
import debias_vision_lang_b200 as dvl
from debias_vision_lang_b200 import _capi, measuring_bias, synthetic, datasets


def test_gen_prompts_equals_reference_output():
    gold = load_golden("prompts.json")
    prompts = dvl.gen_prompts()
    assert len(prompts) == gold["n"] == 319
    assert prompts == gold["prompts"]


def test_header_symbols_are_exported():
    header = open(os.path.join(ROOT, "include", "dvl_b200.h")).read()
    declared = set(re.findall(r"DVL_API\s+[\w\s\*]+?\b(dvl_\w+)\s*\(", header))
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    lib = _capi.lib()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.dvl_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    with pytest.raises(_capi.DvlError):
        _capi.Context(0)
    w = synthetic.dyadic_grid(64, 16, 2, 2)
    with pytest.raises(RuntimeError):
        dvl.eval_ranking(w.labels, w.gallery, w.queries, "maxskew", 10)
    import pandas as pd
    with pytest.raises(RuntimeError):
        dvl.compute_skew_metrics(pd.DataFrame({"score": [0.1, 0.2], "label": [0, 1]}), 1)


def test_label_encoding_and_errors():
    codes, c = measuring_bias._encode_labels(np.array([0, 1, 1, 0, 2]), True)
    assert c == 3 and codes.dtype == np.uint8 and codes.tolist() == [0, 1, 1, 0, 2]
    # non-contiguous labels: fine for maxskew (Counter keys), IndexError for ndkl like the reference
    codes, c = measuring_bias._encode_labels(np.array([0, 2, 2, 0]), False)
    assert c == 2 and codes.tolist() == [0, 1, 1, 0]
    with pytest.raises(IndexError):
        measuring_bias._encode_labels(np.array([0, 2, 2, 0]), True)
    with pytest.raises(ValueError):
        measuring_bias._encode_labels(np.arange(40), False)


def test_topn_resolution_matches_reference_rule():
    assert measuring_bias._resolve_topn(1.0, 10954) == 10954
    assert measuring_bias._resolve_topn(0.1, 10954) == 1096
    assert measuring_bias._resolve_topn(1000, 5) == 1000
    with pytest.raises(ZeroDivisionError):
        measuring_bias._resolve_topn(0, 5)


def test_eval_ranking_asserts_on_bad_evaluation():
    w = synthetic.dyadic_grid(64, 16, 2, 2)
    with pytest.raises(AssertionError):
        dvl.eval_ranking(w.labels, w.gallery, w.queries, "skew", 10)


def test_dotdict_and_encodings():
    d = dvl.Dotdict(a=1)
    assert d.a == 1 and d.missing is None
    d.b = 2
    assert d["b"] == 2
    assert datasets.IATDataset.GENDER_ENCODING == {"Female": 1, "Male": 0}
    assert len(datasets.IATDataset.AGE_ENCODING) == 9 and len(datasets.FairFace.RACE_ENCODING) == 7


def test_fairface_mirror_on_a_tiny_directory(tmp_path):
    import pandas as pd
    root = tmp_path / "fairface"
    (root / "labels" / "val").mkdir(parents=True)
    (root / "imgs" / "train_val" / "val").mkdir(parents=True)
    rows = []
    races = list(datasets.FairFace.RACE_ENCODING)
    for i in range(1, 29):
        rows.append(dict(file=f"val/{i}.jpg", age="20-29", gender="Male" if i % 3 else "Female",
                         race=races[i % 7], service_test=False))
    pd.DataFrame(rows).to_csv(root / "labels" / "val" / "val_labels.csv", index=False)
    ds = datasets.FairFace(iat_type="race", mode="val", equal_split=False, data_path=str(root))
    assert len(ds) == 28 and ds.n_iat_classes == 7
    # lexicographic file order (val/1, val/10, val/11, ...) defines the gallery index order
    assert list(ds.labels["file"][:3]) == ["val/1.jpg", "val/10.jpg", "val/11.jpg"]
    ds2 = datasets.FairFace(iat_type="gender", mode="val", equal_split=True, data_path=str(root))
    g = list(ds2.labels["gender"])
    assert g[:len(g) // 2] == ["Male"] * (len(g) // 2) and g[len(g) // 2:] == ["Female"] * (len(g) // 2)
    with pytest.raises(FileNotFoundError):
        datasets.FairFace(iat_type="gender", mode="val", data_path=str(tmp_path / "nope"))


def test_synthetic_workloads_are_deterministic_and_exact():
    a = synthetic.dyadic_grid(128, 64, 4, 7, seed=3)
    b = synthetic.dyadic_grid(128, 64, 4, 7, seed=3)
    assert torch.equal(a.gallery, b.gallery) and np.array_equal(a.labels, b.labels)
    assert set(np.unique(a.labels)) == set(range(7))
    s64 = a.queries.double() @ a.gallery.double().T
    s32 = a.queries @ a.gallery.T
    assert torch.equal(s32.double(), s64)          # exactly representable, any summation order
    adv = synthetic.adversarial(512, 64, 16, 2, seed=6)
    assert float(adv.queries[0].abs().sum()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_gallery_cache_round_trip_and_shards(tmp_path, dtype):
    """EmbeddingGallery.save/load (the `use_cache` the reference reserves, datasets.py:26): bit-exact
    round trip in every stored dtype, and rank shards that tile the gallery in order."""
    from debias_vision_lang_b200.sharded import shard_bounds
    w = synthetic.clip_like(1300, 24, 3, 7, seed=3)
    emb = w.gallery.to(dtype)
    path = str(tmp_path / "gallery")
    assert not datasets.EmbeddingGallery.is_cached(path)
    datasets.EmbeddingGallery(emb, w.labels, "race").save(path)
    assert datasets.EmbeddingGallery.is_cached(path)
    whole = datasets.EmbeddingGallery.load(path)
    assert whole.image_embeddings.dtype == dtype and torch.equal(whole.image_embeddings, emb)
    assert np.array_equal(whole.iat_labels, w.labels) and whole.n_iat_classes == 7 and whole.iat_type == "race"
    assert whole.index_base == 0 and whole.n_global == 1300
    rows, labs = [], []
    for world in (3,):
        for r in range(world):
            part = datasets.EmbeddingGallery.load(path, rank=r, world=world)
            lo, hi = shard_bounds(1300, world, r)
            assert part.index_base == lo and len(part) == hi - lo and part.n_global == 1300
            assert part.n_iat_classes == 7          # a shard may miss a class; the count is the gallery's
            rows.append(part.image_embeddings)
            labs.append(part.iat_labels)
            with pytest.raises(ValueError):
                part.save(str(tmp_path / "shard")) if len(part) != 1300 else (_ for _ in ()).throw(ValueError())
    assert torch.equal(torch.cat(rows), emb) and np.array_equal(np.concatenate(labs), w.labels)
    item = whole[5]
    assert torch.equal(item["img"], emb[5]) and item["iat_label"] == int(w.labels[5])


def test_float64_scores_get_an_order_and_tie_preserving_float32_surrogate():
    from debias_vision_lang_b200.measuring_bias import _scores_as_f32
    exact = np.array([0.5, 0.25, -1.0, 3.0], dtype=np.float64)
    assert _scores_as_f32(exact).dtype == np.float32 and np.array_equal(_scores_as_f32(exact), exact.astype(np.float32))
    close = np.array([1.0, 1.0 + 1e-12, 1.0, 1.0 - 1e-12, np.nan, 2.0])
    s = _scores_as_f32(close)
    assert s[1] > s[0] == s[2] > s[3] and s[5] > s[1] and np.isnan(s[4])
    assert np.array_equal(_scores_as_f32(np.array([3, 1, 2])), np.array([3, 1, 2], dtype=np.float32))


def test_header_is_plain_c_and_the_library_binds_from_c(tmp_path):
    """include/dvl_b200.h compiles as C99 (no C++, no torch types) and a C program can bind the library the way
    any FFI would: dlopen, version, context creation (DVL_ERR_NO_DEVICE without a B200 -- never a silent fallback)."""
    import shutil
    import subprocess
    from conftest import ROOT
    from debias_vision_lang_b200 import _capi
    if shutil.which("gcc") is None or not os.path.exists(_capi.LIB_PATH):
        pytest.skip("needs gcc and the built library")
    src = tmp_path / "bind.c"
    src.write_text(r'''
#include <dlfcn.h>
#include <stdio.h>
#include "dvl_b200.h"
typedef int (*version_fn)(void);
typedef int (*create_fn)(dvl_ctx**, int);
typedef int (*destroy_fn)(dvl_ctx*);
int main(int argc, char** argv) {
  void* lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { printf("dlopen failed: %s\n", dlerror()); return 2; }
  version_fn version; create_fn create; destroy_fn destroy;
  *(void**)(&version) = dlsym(lib, "dvl_abi_version");     /* the POSIX idiom: ISO C has no object -> function cast */
  *(void**)(&create) = dlsym(lib, "dvl_create");
  *(void**)(&destroy) = dlsym(lib, "dvl_destroy");
  if (!version || !create || !destroy || !dlsym(lib, "dvl_eval_ranking") || !dlsym(lib, "dvl_shard_topk_host")) return 3;
  dvl_ctx* ctx = 0;
  int rc = create(&ctx, 0);
  printf("abi=%d create=%d\n", version(), rc);
  if (rc == DVL_OK) destroy(ctx);
  return (version() == DVL_B200_ABI_VERSION && (rc == DVL_OK || rc == DVL_ERR_NO_DEVICE)) ? 0 : 4;
}
''')
    exe = tmp_path / "bind"
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-ldl"])
    out = subprocess.run([str(exe), _capi.LIB_PATH], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi=1" in out.stdout

"""Dataset interface of the reference (``debias_clip/datasets.py``) as the hot path sees it:
items are dict-likes with ``"img"`` and ``"iat_label"`` (datasets.py:125-131), attribute
strings map to contiguous int labels 0..C-1 (datasets.py:17-19,50-51, gen_labels :30-46).

This side of the boundary is I/O, not the accelerated path.  What is here: the label
encodings, ``gen_labels`` with the reference's checks, a ``FairFace`` that reads an already
present FairFace directory with the reference's ordering rules but runs on pandas >= 2
(the reference's ``equal_split`` crashes there) and never downloads (no network),
and ``EmbeddingGallery`` -- the device-resident (embeddings, labels) pair the path consumes.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Union

import numpy as np
import torch
from torch.utils.data import Dataset

from . import Dotdict, FAIRFACE_DATA_PATH


class IATDataset(Dataset):
    GENDER_ENCODING = {"Female": 1, "Male": 0}
    AGE_ENCODING = {"0-2": 0, "3-9": 1, "10-19": 2, "20-29": 3, "30-39": 4,
                    "40-49": 5, "50-59": 6, "60-69": 7, "more than 70": 8}

    def __init__(self):
        self.image_embeddings: Optional[torch.Tensor] = None
        self.iat_labels: Optional[np.ndarray] = None
        self._img_fnames = None
        self._transforms = None
        self.use_cache = None
        self.iat_type = None
        self.n_iat_classes = None

    def gen_labels(self, iat_type: str, label_encoding: Optional[dict] = None):
        """attribute name -> (int labels, number of classes); every class must occur
        (datasets.py:43) -- the metrics index their count arrays with the label."""
        if iat_type in ("gender_science", "test_weat", "gender"):
            column, encoding = "gender", IATDataset.GENDER_ENCODING
        elif iat_type == "race":
            column, encoding = "race", self.RACE_ENCODING
        elif iat_type == "age":
            column, encoding = "age", IATDataset.AGE_ENCODING
        else:
            raise NotImplementedError
        if label_encoding is not None:
            encoding = label_encoding
        labels_list = self.labels[column]
        assert set(labels_list.unique()) == set(encoding.keys()), "There is a missing label, invalid for WEAT"
        return np.array([encoding[x] for x in labels_list], dtype=int), len(encoding)


class FairFace(IATDataset):
    RACE_ENCODING = {"White": 0, "Southeast Asian": 1, "Middle Eastern": 2,
                     "Black": 3, "Indian": 4, "Latino_Hispanic": 5, "East Asian": 6}

    def __init__(self, iat_type: str = None, lazy: bool = True, mode: str = "train",
                 _n_samples: Union[float, int] = None, transforms: Callable = None, equal_split: bool = True,
                 data_path: Optional[str] = None):
        super().__init__()
        import pandas as pd
        self.DATA_PATH = str(data_path or FAIRFACE_DATA_PATH)
        self.download_data()
        self.mode = mode
        self._transforms = (lambda x: x) if transforms is None else transforms
        self.labels = pd.read_csv(os.path.join(self.DATA_PATH, "labels", mode, f"{mode}_labels.csv"))
        # lexicographic file order defines the gallery index order, hence the tie-break
        self.labels = self.labels.sort_values("file")
        if _n_samples is not None:
            if isinstance(_n_samples, float):
                _n_samples = int(len(self.labels) * _n_samples)
            self.labels = self.labels[:_n_samples]
        if equal_split:
            # datasets.py:66-78: same number of Male and Female rows, all Male rows first
            male = self.labels.loc[self.labels["gender"] == "Male"]
            female = self.labels.loc[self.labels["gender"] == "Female"]
            k = min(len(male), len(female))
            self.labels = pd.concat([male.sample(n=k, random_state=1), female.sample(n=k, random_state=1)],
                                    ignore_index=True)
        self._img_fnames = [os.path.join(self.DATA_PATH, "imgs", "train_val", x) for x in self.labels["file"]]
        self._fname_to_inx = {fname: inx for inx, fname in enumerate(self._img_fnames)}
        self.images_list = None
        if not lazy:
            self.images_list = [self[i] for i in range(len(self.labels))]
        self.iat_type = iat_type
        self.iat_labels, self.n_iat_classes = self.gen_labels(iat_type=iat_type)

    def download_data(self):
        """The reference downloads from Google Drive (datasets.py:89-117); this build has no network,
        so the data must already be in place."""
        need = [os.path.join(self.DATA_PATH, "labels"), os.path.join(self.DATA_PATH, "imgs")]
        missing = [p for p in need if not os.path.isdir(p)]
        if missing:
            raise FileNotFoundError(
                f"FairFace data not found under {self.DATA_PATH} (missing {missing}); place labels/<mode>/"
                "<mode>_labels.csv and imgs/train_val/ there -- downloading is not supported offline")

    def _load_fairface_sample(self, sample_labels) -> dict:
        from PIL import Image
        res = Dotdict(dict(sample_labels))
        res.img = self._transforms(Image.open(os.path.join(self.DATA_PATH, "imgs", "train_val", res.file)))
        return res

    def __getitem__(self, index: int):
        if self.images_list is not None:
            return self.images_list[index]
        sample = self._load_fairface_sample(self.labels.iloc[index])
        sample.iat_label = self.iat_labels[index]
        return sample

    def __len__(self):
        return len(self.labels)


class EmbeddingGallery(IATDataset):
    """A gallery that already is embeddings: items are ``{"img": embedding, "iat_label": label}`` so
    it can flow through ``get_labels_img_embeddings`` with an identity ``encode_image``; the pair
    (``image_embeddings``, ``iat_labels``) fills the two attributes the reference reserves but never
    implements (datasets.py:22-23).

    ``save`` / ``load`` are the on-disk form of that pair -- the ``use_cache`` the reference reserves
    (datasets.py:26) -- and the wire format that feeds gallery shards: a directory with

    * ``embeddings.npy``  [N, D] row-major, float32 / float16, or uint16 holding bfloat16 bit patterns;
    * ``labels.npy``      [N] uint8 class codes 0..C-1, in gallery (tie-break) order;
    * ``meta.json``       format tag, N, D, dtype, iat_type, n_classes.

    ``load(rank, world)`` memory-maps the files and reads only that rank's contiguous row block
    (``sharded.shard_bounds``: 256-row aligned), so a 10 M-row gallery never passes through one host
    buffer; ``index_base`` / ``n_global`` are what ``sharded.evaluate_sharded`` needs."""

    FORMAT = "dvl-gallery/1"

    def __init__(self, image_embeddings: torch.Tensor, iat_labels: np.ndarray, iat_type: str = "gender",
                 index_base: int = 0, n_global: Optional[int] = None, n_classes: Optional[int] = None,
                 tag: Optional[str] = None):
        super().__init__()
        self.tag = tag
        assert len(image_embeddings) == len(iat_labels)
        self.image_embeddings = image_embeddings
        self.iat_labels = np.asarray(iat_labels)
        self.iat_type = iat_type
        self.n_iat_classes = int(n_classes) if n_classes is not None else int(len(np.unique(self.iat_labels)))
        self.index_base = int(index_base)
        self.n_global = int(n_global) if n_global is not None else len(self.iat_labels)

    def __getitem__(self, index: int):
        return {"img": self.image_embeddings[index], "iat_label": int(self.iat_labels[index])}

    def __len__(self):
        return len(self.iat_labels)

    # ---- on-disk cache ----------------------------------------------------------------------
    def save(self, path: str) -> str:
        """Writes the whole (unsharded) gallery to directory ``path``; returns ``path``."""
        import json
        if self.index_base != 0 or self.n_global != len(self):
            raise ValueError("only a whole gallery can be saved (this object is one shard of it)")
        lab = np.asarray(self.iat_labels)
        if lab.size and (lab.min() < 0 or lab.max() > 255):
            raise ValueError("labels must be class codes 0..255")
        emb = self.image_embeddings.detach().cpu().contiguous()
        if emb.dtype == torch.bfloat16:
            arr, dtype = emb.view(torch.int16).numpy().view(np.uint16), "bfloat16"
        elif emb.dtype in (torch.float16, torch.float32):
            arr, dtype = emb.numpy(), str(emb.dtype).replace("torch.", "")
        else:
            arr, dtype = emb.float().numpy(), "float32"
        os.makedirs(path, exist_ok=True)
        np.save(os.path.join(path, "embeddings.npy"), arr)
        np.save(os.path.join(path, "labels.npy"), lab.astype(np.uint8))
        meta = {"format": self.FORMAT, "n": int(arr.shape[0]), "d": int(arr.shape[1]), "dtype": dtype,
                "iat_type": self.iat_type, "n_classes": int(self.n_iat_classes), "tag": self.tag}
        with open(os.path.join(path, "meta.json"), "w") as f:      # written last: its presence marks a complete cache
            json.dump(meta, f)
        return path

    @staticmethod
    def is_cached(path: str) -> bool:
        return os.path.isfile(os.path.join(path, "meta.json"))

    @classmethod
    def load(cls, path: str, rank: int = 0, world: int = 1, device: Optional[Union[str, torch.device]] = None,
             expect_iat_type: Optional[str] = None, expect_tag: Optional[str] = None) -> "EmbeddingGallery":
        """Rows ``shard_bounds(N, world, rank)`` of the cached gallery, optionally moved to ``device``.
        ``expect_iat_type`` / ``expect_tag``: what the caller is about to measure; a cache written for a
        different attribute or model tag raises ``ValueError`` instead of being silently reused."""
        import json
        from .sharded import shard_bounds
        with open(os.path.join(path, "meta.json")) as f:
            meta = json.load(f)
        if meta.get("format") != cls.FORMAT:
            raise ValueError(f"{path}: not a {cls.FORMAT} gallery cache")
        if expect_iat_type is not None and meta.get("iat_type") != expect_iat_type:
            raise ValueError(f"{path}: cache holds labels of attribute {meta.get('iat_type')!r}, "
                             f"not {expect_iat_type!r}; delete it or point use_cache elsewhere")
        if expect_tag is not None and meta.get("tag") != expect_tag:
            raise ValueError(f"{path}: cache was written for model tag {meta.get('tag')!r}, not {expect_tag!r}")
        n = int(meta["n"])
        lo, hi = shard_bounds(n, world, rank)
        arr = np.load(os.path.join(path, "embeddings.npy"), mmap_mode="r")
        lab = np.load(os.path.join(path, "labels.npy"), mmap_mode="r")
        if arr.shape != (n, int(meta["d"])) or lab.shape != (n,):
            raise ValueError(f"{path}: array shapes do not match meta.json")
        block = np.array(arr[lo:hi])                  # a private, writable copy of this rank's rows
        if meta["dtype"] == "bfloat16":
            emb = torch.from_numpy(block.view(np.int16)).view(torch.bfloat16)
        else:
            emb = torch.from_numpy(block)
        if device is not None:
            emb = emb.to(device)
        lab_block = np.array(lab[lo:hi]).astype(np.int64)
        if lab_block.size and int(lab_block.max()) >= int(meta["n_classes"]):
            raise ValueError(f"{path}: a label code is >= n_classes ({meta['n_classes']})")
        return cls(emb, lab_block, meta.get("iat_type", "gender"),
                   index_base=lo, n_global=n, n_classes=int(meta["n_classes"]), tag=meta.get("tag"))

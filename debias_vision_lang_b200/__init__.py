"""B200-native drop-in for the bias-measurement hot path of oxai/debias-vision-lang
(``debias_clip.measure_bias`` after the encoders).  Mirrors ``debias_clip/__init__.py:1-26``
for this path: ``Dotdict``, the data-path constants and ``measure_bias``; the arithmetic
runs in hand-written sm_100a CUDA behind ``include/dvl_b200.h`` (no CPU fallback)."""
import pathlib
from typing import Any


class Dotdict(dict):
    """Attribute-style dict (debias_clip/__init__.py:4-18); missing keys read as None."""

    def __getattr__(self, name: str) -> Any:
        return super().get(name)

    def __setattr__(self, name: str, value: Any) -> None:
        return super().__setitem__(name, value)

    def __delattr__(self, name: str) -> None:
        return super().__delitem__(name)

    def __getstate__(self):
        return self.__dict__

    def __setstate__(self, d):
        self.__dict__.update(d)


DATA_PATH = (pathlib.Path(__file__).parent / "data").resolve()
FAIRFACE_DATA_PATH = DATA_PATH / "fairface"
PROMPT_DATA_PATH = DATA_PATH / "prompt_templates.csv"   # not read: the prompt array is built in

from .measuring_bias import (  # noqa: E402
    DEFAULT_OPTS, adversary_input, clip_logits, compute_skew_metrics, eval_ranking, evaluate, gen_prompts, get_labels_img_embeddings,
    get_prompt_embeddings, measure_bias, normalized_discounted_KL,
)

__all__ = [
    "Dotdict", "DATA_PATH", "FAIRFACE_DATA_PATH", "PROMPT_DATA_PATH", "measure_bias", "eval_ranking", "evaluate",
    "compute_skew_metrics", "normalized_discounted_KL", "gen_prompts", "get_labels_img_embeddings",
    "get_prompt_embeddings", "DEFAULT_OPTS", "clip_logits", "adversary_input",
]

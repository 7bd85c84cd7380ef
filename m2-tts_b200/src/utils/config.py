"""YAML config loading with the attribute/.get access the reference gets from OmegaConf
(reference scripts/synthesize.py:37-46, training/train.py:157-166)."""
from __future__ import annotations

from pathlib import Path
from typing import Any, Union

import yaml


class AttrDict(dict):
    """dict with attribute access, nested; `.get(key, default)` works as on a DictConfig."""

    def __getattr__(self, key: str) -> Any:
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    @classmethod
    def wrap(cls, obj: Any) -> Any:
        if isinstance(obj, dict):
            return cls({k: cls.wrap(v) for k, v in obj.items()})
        if isinstance(obj, list):
            return [cls.wrap(v) for v in obj]
        return obj


def load_config(path: Union[str, Path]) -> AttrDict:
    with open(path, "r") as f:
        return AttrDict.wrap(yaml.safe_load(f))


def model_kwargs(config: Any) -> dict:
    """The constructor arguments scripts/synthesize.py:37-46 derives from a config."""
    m = config.model
    return dict(vocab_size=m.text_encoder.vocab_size, hidden_dim=m.text_encoder.hidden_dim,
                mel_channels=m.decoder.mel_channels, text_encoder_layers=m.text_encoder.num_layers,
                decoder_layers=m.decoder.get("num_layers", 2), num_heads=m.text_encoder.num_heads,
                dropout=m.text_encoder.dropout, vocoder_channels=m.vocoder.hidden_channels)

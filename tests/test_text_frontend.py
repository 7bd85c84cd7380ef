"""Text front-end parity: our batched TextProcessor against golden vectors generated from
the unmodified reference (tests/golden/make_text_golden.py) and, when /root/reference exists, the live reference."""
import importlib.util
import json
from pathlib import Path

import pytest
import torch

import helpers as H  # noqa: F401  (puts m2-tts_b200/src on sys.path)
from utils.text import PHONEME_SET, TextProcessor

GOLD = json.loads((Path(__file__).parent / "golden" / "text_golden.json").read_text())


def test_symbol_inventory_matches_reference():
    assert PHONEME_SET == GOLD["symbols"]


def test_process_text_matches_golden():
    tp = TextProcessor()
    for c in GOLD["cases"]:
        r = tp.process_text(c["text"], c["max_length"])
        assert r["phoneme_ids"] == c["phoneme_ids"], c["text"]
        assert r["length"] == c["length"], c["text"]
        assert r["phonemes"] == tp.ids_to_phonemes(r["phoneme_ids"])


def test_hello_world_is_c1():
    r = TextProcessor().process_text("Hello world", max_length=256)
    assert r["phoneme_ids"][:11] == [39, 21, 6, 24, 11, 40, 35, 7, 24, 17, 39] and r["length"] == 9
    assert set(r["phoneme_ids"][11:]) == {39} and len(r["phoneme_ids"]) == 256


def test_process_batch_equals_per_sentence():
    tp = TextProcessor()
    texts = [c["text"] for c in GOLD["cases"] if c["max_length"] == 256][:40]
    ids, lengths = tp.process_batch(texts, max_length=256)
    assert ids.dtype == torch.int64 and ids.shape == (len(texts), 256)
    for b, t in enumerate(texts):
        r = tp.process_text(t, 256)
        assert ids[b].tolist() == r["phoneme_ids"] and int(lengths[b]) == r["length"]
    ids2, _ = tp.process_batch(["hello", "hello world and more"])
    assert ids2.shape[1] == len(tp.process_text("hello world and more")["phoneme_ids"])


@pytest.mark.skipif(not Path("/root/reference/src/utils/text.py").exists(), reason="reference not mounted")
def test_against_live_reference():
    spec = importlib.util.spec_from_file_location("reftext", "/root/reference/src/utils/text.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    a, b = ref.TextProcessor(), TextProcessor()
    for s in ["We can find a long way down", "Numbers 1 2 3 19 20 21", "it's  spaced\tout\nhere", "x", "!?"]:
        for ml in (None, 8, 64):
            assert a.process_text(s, ml) == b.process_text(s, ml)

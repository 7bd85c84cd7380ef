"""The product-side stage configurations (models/stage_configs.py) are the ones the oracle — i.e. the reference's YAML files
as scripts/synthesize.py consumes them — uses."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "m2-tts_b200" / "src")):
    if p not in sys.path:
        sys.path.insert(0, p)

from models import stage_configs  # noqa: E402
from oracle import m2tts_oracle as oracle  # noqa: E402


def test_stage_kwargs_match_the_oracle():
    assert stage_configs.STAGE_KWARGS == oracle.STAGE_KWARGS
    assert stage_configs.SAMPLES_PER_FRAME == oracle.SAMPLES_PER_FRAME
    assert stage_configs.SAMPLE_RATE == oracle.SAMPLE_RATE

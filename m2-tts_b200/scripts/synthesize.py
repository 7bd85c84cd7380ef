#!/usr/bin/env python3
"""Text -> wav on a B200: same command line as the reference's scripts/synthesize.py:88-155
(--text --checkpoint --output --duration-scale --sample-rate), plus --texts-file for batched synthesis
(one utterance per line, written as <output stem>_<n>.wav) and --config to supply the YAML when the checkpoint
carries none. The model runs through the C-ABI CUDA library; there is no CPU fallback."""
import argparse
import logging
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "src"))

from utils.audio import save_audio  # noqa: E402
from utils.checkpoint import load_model  # noqa: E402
from utils.config import load_config  # noqa: E402
from utils.device import setup_device  # noqa: E402
from utils.text import TextProcessor  # noqa: E402

logging.basicConfig(level=logging.INFO)
logger = logging.getLogger("synthesize")


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="M2 TTS text synthesis (B200)")
    ap.add_argument("--text", type=str, help="Text to synthesize")
    ap.add_argument("--texts-file", type=str, help="File with one utterance per line (batched synthesis)")
    ap.add_argument("--checkpoint", type=str, required=True, help="Path to model checkpoint")
    ap.add_argument("--config", type=str, default=None, help="YAML config (overrides the checkpoint's)")
    ap.add_argument("--output", type=str, default="output.wav", help="Output audio file path")
    ap.add_argument("--duration-scale", type=float, default=1.0, help="Duration scaling factor (1.0 = normal speed)")
    ap.add_argument("--sample-rate", type=int, default=22050, help="Audio sample rate")
    args = ap.parse_args(argv)
    if (args.text is None) == (args.texts_file is None):
        ap.error("give exactly one of --text / --texts-file")

    device = setup_device()
    model, ckpt = load_model(Path(args.checkpoint), device, load_config(args.config) if args.config else None)
    logger.info("Loaded model from %s (training step: %s)", args.checkpoint, ckpt.get("step", "unknown"))

    texts = [args.text] if args.text is not None else [l.strip() for l in Path(args.texts_file).read_text().splitlines() if l.strip()]
    ids, lengths = TextProcessor().process_batch(texts, max_length=256, pin_memory=True)
    with torch.no_grad():
        mel, audio = model.inference(ids.to(device, non_blocking=True), lengths.to(device, non_blocking=True),
                                     duration_scale=args.duration_scale)
    logger.info("Generated mel %s, audio %s", tuple(mel.shape), tuple(audio.shape))
    if audio is None or audio.size(0) == 0:
        logger.error("No audio generated")
        return 1
    out = Path(args.output)
    out.parent.mkdir(parents=True, exist_ok=True)
    # valid samples per utterance = 64 x frames (rows beyond an utterance's frame count are padding)
    frames = getattr(model.length_regulator, "last_frames", None)
    for n in range(len(texts)):
        wav = audio[n, 0]
        if frames is not None:
            wav = wav[: max(1, int(frames[n])) * 64]
        path = out if len(texts) == 1 else out.with_name(f"{out.stem}_{n}{out.suffix}")
        save_audio(wav, path, args.sample_rate)
        logger.info("Audio saved to: %s (%.2f s)", path, wav.numel() / args.sample_rate)
    return 0


if __name__ == "__main__":
    sys.exit(main())

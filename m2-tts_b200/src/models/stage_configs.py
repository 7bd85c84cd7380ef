"""Constructor arguments of `M2TTSModel` for the reference's shipped configurations.

configs/stage1_poc.yaml:6-27 and configs/stage2_quality.yaml:6-28, as scripts/synthesize.py:37-46 turns them into
`M2TTSModel(...)` keyword arguments; "tiny" is the model of scripts/test_pipeline.py:72-80. Product-side copy: `bench.py`,
`__graft_entry__.py` and the tools build their models from here, not from the test oracle.
"""

STAGE_KWARGS = {
    "stage1": dict(vocab_size=256, hidden_dim=64, mel_channels=64, text_encoder_layers=2,
                   decoder_layers=2, num_heads=2, dropout=0.1, vocoder_channels=128),
    "stage2": dict(vocab_size=256, hidden_dim=96, mel_channels=80, text_encoder_layers=3,
                   decoder_layers=3, num_heads=2, dropout=0.1, vocoder_channels=256),
    "tiny": dict(vocab_size=256, hidden_dim=32, mel_channels=32, text_encoder_layers=1,
                 decoder_layers=1, num_heads=2, dropout=0.1, vocoder_channels=64),
}

SAMPLES_PER_FRAME = 64      # prod([4, 4, 2, 2]), src/models/tts_model.py:244
SAMPLE_RATE = 22050         # configs/stage2_quality.yaml:69

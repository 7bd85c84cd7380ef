"""Generates tests/golden/text_golden.json by importing the UNMODIFIED reference text front-end
(/root/reference/src/utils/text.py) in the build container. Run: python tests/golden/make_text_golden.py"""
import importlib.util
import json
import random
import string
from pathlib import Path

spec = importlib.util.spec_from_file_location("reftext", "/root/reference/src/utils/text.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
tp = ref.TextProcessor()

sentences = ["Hello world", "Dr. Smith has 3 cats & 12 dogs, etc.", "", "   ", "The quick brown fox jumps over the lazy dog!",
             "e.g. this is St. Mary's vs. Mr. O'Neil at 20:30", "Ünïcödé café naïve 5.", "...",
             "hello,world (hello) 'world'", "21 10 7, 0!", "Mrs. and Ms. i.e. that", "a" * 300,
             "They were about to find which way would make more time for him"]
rng = random.Random(0)
alphabet = string.ascii_letters + string.digits + ".,!?&' "
for _ in range(60):
    sentences.append(" ".join("".join(rng.choice(alphabet) for _ in range(rng.randint(1, 9))) for _ in range(rng.randint(0, 12))))
cases = []
for s in sentences:
    for ml in (None, 16, 256):
        r = tp.process_text(s, ml)
        cases.append({"text": s, "max_length": ml, "phoneme_ids": r["phoneme_ids"], "length": r["length"]})
Path(__file__).with_name("text_golden.json").write_text(json.dumps({"symbols": ref.PHONEME_SET, "cases": cases}))
print(len(cases), "cases")

"""Batched text front-end: text -> phoneme ids, same ids and lengths as the reference's
`src/utils/text.py` (`TextProcessor.process_text`, :245-347; `SimpleG2P.convert`, :213-243; `normalize_text`, :84-101)
so `scripts/synthesize.py:66-71` runs unchanged — plus `process_batch`, which returns the padded `[B, S]` id tensor
and `[B]` lengths the model takes (at >= 1000x real time the per-utterance Python loop of the
reference becomes the bottleneck for short utterances; here words are resolved once through a cache).

Restated behaviour (reference file:line):
  * lower-case, NFD-normalise, expand abbreviations by plain substring replacement in table order, spell out the
    numbers 0..20 when a whitespace token is exactly that number (punctuation around it preserved), squeeze whitespace
    (:30-101);
  * per whitespace token: strip leading/trailing ASCII punctuation, look the word up in the lexicon, else map letters
    one by one (consonants fixed, vowels a/e/i/o/u -> AE/EH/IH/AO/UH, everything else skipped, nothing left -> UNK)
    (:184-211); `SP` between words, `SIL` at both ends (:226-243);
  * ids = index in the 42-symbol inventory (:14-27); pad with `SIL` / truncate to `max_length`; `length` counts the
    symbols that are not `SIL` after padding/truncation (:322-347) — so the leading/trailing `SIL` are NOT counted.
"""
from __future__ import annotations

import re
import string
import unicodedata
from typing import Dict, List, Optional, Sequence, Tuple

SYMBOLS: Tuple[str, ...] = tuple(
    "AA AE AH AO AW AY EH ER EY IH IY OW OY UH UW "
    "B CH D DH F G HH JH K L M N NG P R S SH T TH V W Y Z ZH SIL SP UNK".split())
PHONEME_SET = list(SYMBOLS)
PHONEME_TO_ID: Dict[str, int] = {s: i for i, s in enumerate(SYMBOLS)}
ID_TO_PHONEME: Dict[int, str] = dict(enumerate(SYMBOLS))
SIL, SP, UNK = PHONEME_TO_ID["SIL"], PHONEME_TO_ID["SP"], PHONEME_TO_ID["UNK"]

_ABBREVIATIONS = (("dr.", "doctor"), ("mr.", "mister"), ("mrs.", "missus"), ("ms.", "miss"), ("st.", "saint"),
                  ("etc.", "et cetera"), ("vs.", "versus"), ("e.g.", "for example"), ("i.e.", "that is"), ("&", "and"))
_NUMBERS = ("zero one two three four five six seven eight nine ten eleven twelve thirteen fourteen fifteen sixteen "
            "seventeen eighteen nineteen twenty").split()
_LETTERS = dict(zip("bcdfghjklmnpqrstvwxyz", "B K D F G HH JH K L M N P K R S T V W K Y Z".split()))
_LETTERS.update(zip("aeiou", "AE EH IH AO UH".split()))
_LEXICON_DATA = (
    "hello=HH EH L OW;world=W ER L D;the=DH AH;and=AE N D;to=T UW;a=AH;of=AH V;in=IH N;is=IH Z;it=IH T;you=Y UW;"
    "that=DH AE T;he=HH IY;was=W AH Z;for=F ER;on=AO N;are=AA R;as=AE Z;with=W IH TH;his=HH IH Z;they=DH EY;i=AY;"
    "at=AE T;be=B IY;this=DH IH S;have=HH AE V;from=F R AH M;or=ER;one=W AH N;had=HH AE D;by=B AY;word=W ER D;"
    "but=B AH T;not=N AA T;what=W AH T;all=AO L;were=W ER;we=W IY;when=W EH N;your=Y ER;can=K AE N;said=S EH D;"
    "there=DH EH R;each=IY CH;which=W IH CH;do=D UW;how=HH AW;their=DH EH R;if=IH F;will=W IH L;up=AH P;"
    "other=AH DH ER;about=AH B AW T;out=AW T;many=M EH N IY;then=DH EH N;them=DH EH M;these=DH IY Z;so=S OW;"
    "some=S AH M;her=HH ER;would=W UH D;make=M EY K;like=L AY K;into=IH N T UW;him=HH IH M;time=T AY M;two=T UW;"
    "more=M ER;go=G OW;no=N OW;way=W EY;could=K UH D;my=M AY;than=DH AE N;first=F ER S T;been=B IH N;call=K AO L;"
    "who=HH UW;its=IH T S;now=N AW;find=F AY N D;long=L AO NG;down=D AW N;day=D EY;did=D IH D;get=G EH T;"
    "come=K AH M;made=M EY D;may=M EY;part=P AA R T"
)
_PUNCT = string.punctuation
_WS = re.compile(r"\s+")


def _lexicon() -> Dict[str, Tuple[int, ...]]:
    out = {}
    for entry in _LEXICON_DATA.split(";"):
        word, phones = entry.split("=")
        out[word] = tuple(PHONEME_TO_ID[p] for p in phones.split())
    return out


def normalize_text(text: str) -> str:
    text = unicodedata.normalize("NFD", text.lower())
    for short, full in _ABBREVIATIONS:
        text = text.replace(short, full)
    tokens = []
    for tok in text.split():
        core = tok.strip(_PUNCT)
        if core.isdigit() and core in _NUMBER_WORDS:
            lead = tok[:len(tok) - len(tok.lstrip(_PUNCT))]
            trail = tok[len(tok.rstrip(_PUNCT)):]
            tok = lead + _NUMBER_WORDS[core] + trail
        tokens.append(tok)
    return _WS.sub(" ", " ".join(tokens).strip())


_NUMBER_WORDS = {str(i): w for i, w in enumerate(_NUMBERS)}


class TextProcessor:
    """Drop-in for the reference class of the same name, with a batched entry point."""

    def __init__(self, vocab_size: int = 256):
        self.vocab_size = vocab_size
        self.phoneme_to_id = PHONEME_TO_ID
        self.id_to_phoneme = ID_TO_PHONEME
        self._words: Dict[str, Tuple[int, ...]] = _lexicon()      # grows: spelled-out words are cached

    # ---- word level -------------------------------------------------------------------------
    def _word_ids(self, token: str) -> Tuple[int, ...]:
        word = token.strip(_PUNCT)
        ids = self._words.get(word)
        if ids is None:
            ids = tuple(PHONEME_TO_ID[_LETTERS[ch]] for ch in word.lower() if ch in _LETTERS) or (UNK,)
            self._words[word] = ids
        return ids

    def _sentence_ids(self, text: str) -> List[int]:
        ids: List[int] = [SIL]
        first = True
        for token in normalize_text(text).split():
            if not first:
                ids.append(SP)
            ids.extend(self._word_ids(token))
            first = False
        ids.append(SIL)
        return ids

    # ---- reference API ----------------------------------------------------------------------
    def text_to_phonemes(self, text: str) -> List[str]:
        return [SYMBOLS[i] for i in self._sentence_ids(text)]

    def phonemes_to_ids(self, phonemes: Sequence[str]) -> List[int]:
        return [PHONEME_TO_ID.get(p, UNK) for p in phonemes]

    def ids_to_phonemes(self, ids: Sequence[int]) -> List[str]:
        return [ID_TO_PHONEME.get(int(i), "UNK") for i in ids]

    def process_text(self, text: str, max_length: Optional[int] = None) -> Dict:
        ids = self._sentence_ids(text)
        if max_length is not None:
            ids = ids[:max_length] + [SIL] * max(0, max_length - len(ids))
        return {"text": text, "phonemes": [SYMBOLS[i] for i in ids], "phoneme_ids": ids,
                "length": sum(1 for i in ids if i != SIL)}

    # ---- batched entry point ----------------------------------------------------------------
    def process_batch(self, texts: Sequence[str], max_length: Optional[int] = None, pin_memory: bool = False):
        """-> (ids [B, S] int64, lengths [B] int64) with S = max_length or the longest sequence; same per-row content
        and `length` as `process_text(text, max_length=S)`."""
        import torch
        rows = [self._sentence_ids(t) for t in texts]
        S = max_length if max_length is not None else max((len(r) for r in rows), default=0)
        ids = torch.full((len(rows), S), SIL, dtype=torch.int64, pin_memory=pin_memory)
        lengths = torch.zeros((len(rows),), dtype=torch.int64, pin_memory=pin_memory)
        for b, r in enumerate(rows):
            r = r[:S]
            ids[b, :len(r)] = torch.tensor(r, dtype=torch.int64)
            lengths[b] = sum(1 for i in r if i != SIL)
        return ids, lengths

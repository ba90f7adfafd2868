"""Host-side text helpers around the hot path (outside the timed region).

Restates /root/reference/src/inference.rs:276-313 (`parse_asr_output`, `capitalize_first`) and wraps the
HF `tokenizers` runtime like /root/reference/src/tokenizer.rs:4-50.  SURVEY.md section 8f-3.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple


def capitalize_first(s: str) -> str:                       # inference.rs:307-313
    return s[:1].upper() + s[1:] if s else ""


def parse_asr_output(raw: str, language_forced: bool) -> Tuple[str, str]:
    """(language, text) from the decoded string.  inference.rs:276-305."""
    if language_forced:
        return "forced", raw.strip()
    raw = raw.strip()
    if raw.startswith("language "):
        rest = raw[len("language "):]
        pos = rest.find("<asr_text>")
        if pos >= 0:
            return rest[:pos].strip(), rest[pos + len("<asr_text>"):].strip()
        lang_end = 0
        for i, c in enumerate(rest):
            if c.isspace() or not c.isalpha():
                lang_end = i
                break
            lang_end = i + 1
        if lang_end > 0:
            return rest[:lang_end], rest[lang_end:].strip()
    return "unknown", raw


class AsrTokenizer:
    """tokenizer.json wrapper (tokenizer.rs:4-50): encode without special tokens, decode skipping them."""

    def __init__(self, tok):
        self._tok = tok

    @classmethod
    def from_dir(cls, model_dir: str) -> "AsrTokenizer":
        path = os.path.join(model_dir, "tokenizer.json")
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"tokenizer.json not found in {model_dir}; generate it with transformers.AutoTokenizer(...)"
                ".backend_tokenizer.save(...) as the reference's tokenizer.rs:22-31 instructs")
        from tokenizers import Tokenizer
        return cls(Tokenizer.from_file(path))

    def encode(self, text: str) -> List[int]:
        return list(self._tok.encode(text, add_special_tokens=False).ids)

    def decode(self, ids: List[int]) -> str:
        return self._tok.decode(list(ids), skip_special_tokens=True)


def language_prompt_ids(tokenizer: Optional[AsrTokenizer], language: Optional[str]) -> Optional[List[int]]:
    """ids of "language Xxx" appended to the prompt when the language is forced (inference.rs:246-250)."""
    if language is None:
        return None
    if tokenizer is None:
        raise ValueError("forcing a language needs tokenizer.json (to encode the prompt suffix)")
    return tokenizer.encode(f"language {capitalize_first(language)}")

"""Long-form text splitting: the host-side front of the reference's `read.py` loop.

`split_and_recombine_text` must produce exactly the chunks of the reference's `utils/text.py:4-72` because the
chunk boundaries decide what each `tts()` call synthesises. The reference walks the text with a cursor that can step
backwards; its quirks are part of the observable behaviour and are kept (each is marked below):

  * stepping BACK over a character toggles the quote state on the character the cursor lands on, not on the one it left
    (`utils/text.py:20-30`);
  * look-ahead never sees the last character of the text and returns "" there, and `"" in "..."` is True, so a look-ahead
    past the end counts as "followed by a space" and as "followed by more punctuation" (`utils/text.py:32-34,56,58`).

This is host logic only (no device work); parity is checked against the reference's own known-answer tests
(`utils/text.py:79-103`) and against reference outputs on seeded random texts (`tests/test_text_split.py`).
"""
import re

_BOUNDARY = "!?\n"
_SPACE_LIKE = "\n "
_PUNCT = "!?."


class _Cursor:
    """Position in the normalised text plus the chunk under construction [start, pos]."""

    def __init__(self, text):
        self.text = text
        self.pos = -1
        self.start = 0
        self.in_quote = False
        self.last = len(text) - 1

    def length(self):
        return self.pos - self.start + 1

    def forward(self, n=1):
        for _ in range(n):
            self.pos += 1
            if self.text[self.pos] == '"':
                self.in_quote = not self.in_quote
        return self.text[self.pos]

    def back(self, n=1):
        for _ in range(n):
            self.pos -= 1
            if self.text[self.pos] == '"':          # quirk: toggles on the character landed on
                self.in_quote = not self.in_quote
        return self.text[self.pos]

    def peek(self, delta):
        p = self.pos + delta
        return self.text[p] if 0 <= p < self.last else ""   # quirk: the last character is never visible

    def take(self):
        chunk = self.text[self.start:self.pos + 1]
        self.start = self.pos + 1
        return chunk


def split_and_recombine_text(text, desired_length=200, max_length=300):
    """Chunks of about `desired_length` characters (never more than `max_length`), sentences kept whole when possible.
    Same contract and outputs as the reference function of the same name (`utils/text.py:4-72`)."""
    text = re.sub(r"\n\n+", "\n", text)
    text = re.sub(r"\s+", " ", text)
    text = re.sub(r"[“”]", '"', text)
    cur = _Cursor(text)
    chunks, splits = [], []          # splits: positions inside the current chunk where a sentence ends
    while cur.pos < cur.last:
        c = cur.forward()
        if cur.length() >= max_length:
            # forced split: back to the last sentence end if the chunk is already half full, else to a word boundary
            if splits and cur.length() > desired_length / 2:
                cur.back(cur.pos - splits[-1])
            else:
                while c not in "!?.\n " and cur.pos > 0 and cur.length() > desired_length:
                    c = cur.back()
            chunks.append(cur.take())
            splits = []
        elif not cur.in_quote and (c in _BOUNDARY or (c == "." and cur.peek(1) in _SPACE_LIKE)):
            # sentence end; swallow runs of closing punctuation while they fit
            while cur.pos < len(text) - 1 and cur.length() < max_length and cur.peek(1) in _PUNCT:
                c = cur.forward()
            splits.append(cur.pos)
            if cur.length() >= desired_length:
                chunks.append(cur.take())
                splits = []
        elif cur.in_quote and cur.peek(1) == '"' and cur.peek(2) in _SPACE_LIKE:
            # the end of a quotation followed by white space also ends a sentence
            cur.forward(2)
            splits.append(cur.pos)
    chunks.append(cur.take())
    chunks = [s.strip() for s in chunks]
    return [s for s in chunks if s and not re.match(r"^[\s\.,;:!?]*$", s)]


def utterance_plan(n_chunks, world_size):
    """Which rank renders which chunk of a long text (utterance u -> rank u mod G)."""
    return [u % max(world_size, 1) for u in range(n_chunks)]

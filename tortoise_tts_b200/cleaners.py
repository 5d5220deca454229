"""Text cleaners of the reference tokenizer (tortoise/utils/tokenizer.py:11-157), host side, dependency-free.

`english_cleaners` = unidecode -> lowercase -> expand_numbers -> expand_abbreviations -> collapse whitespace -> strip '"'
(tokenizer.py:146-155); `basic_cleaners` = lowercase -> collapse whitespace (tokenizer.py:129-133).

The reference leans on two packages that are not part of this image: `inflect` (number_to_words) and `unidecode`.
* `number_to_words` below restates the three call forms the reference uses (tokenizer.py:84-103): cardinals with
  `andword=''`, year-style `group=2, zero='oh'`, and ordinals ("243rd" -> "two hundred and forty-third"); pinned by the
  known-answer table of the cleaners' upstream test-suite (tests/test_cleaners.py).
* `to_ascii` covers what unidecode does for Latin-script text (NFKD decomposition, combining marks dropped, a table for
  punctuation / ligatures / currency); characters of other scripts are dropped instead of romanised. ASCII input - the
  only input the English BPE vocabulary was trained on - is returned unchanged. (Deliberate, documented difference.)
"""
import re
import unicodedata

_whitespace_re = re.compile(r"\s+")

_abbreviations = [(re.compile("\\b%s\\." % a, re.IGNORECASE), b) for a, b in [
    ("mrs", "misess"), ("mr", "mister"), ("dr", "doctor"), ("st", "saint"), ("co", "company"), ("jr", "junior"),
    ("maj", "major"), ("gen", "general"), ("drs", "doctors"), ("rev", "reverend"), ("lt", "lieutenant"),
    ("hon", "honorable"), ("sgt", "sergeant"), ("capt", "captain"), ("esq", "esquire"), ("ltd", "limited"),
    ("col", "colonel"), ("ft", "fort")]]          # tokenizer.py:15-34 (spelling 'misess' is the reference's)

_UNITS = ["", "one", "two", "three", "four", "five", "six", "seven", "eight", "nine"]
_TEENS = ["ten", "eleven", "twelve", "thirteen", "fourteen", "fifteen", "sixteen", "seventeen", "eighteen", "nineteen"]
_TENS = ["", "", "twenty", "thirty", "forty", "fifty", "sixty", "seventy", "eighty", "ninety"]
_MILL = ["", " thousand", " million", " billion", " trillion", " quadrillion", " quintillion", " sextillion",
         " septillion", " octillion", " nonillion", " decillion"]
_ORDINAL = {"ty": "tieth", "one": "first", "two": "second", "three": "third", "five": "fifth", "eight": "eighth",
            "nine": "ninth", "twelve": "twelfth"}
_ordinal_suffix_re = re.compile(r"(%s)\Z" % "|".join(_ORDINAL))


def _tenfn(tens, units):
    if tens == 1:
        return _TEENS[units]
    if tens:
        return _TENS[tens] + ("-" + _UNITS[units] if units else "")
    return _UNITS[units]


def _cardinal(num, andword):
    """inflect.number_to_words(num, andword=...) for group=0: 3-digit groups joined by ', '; the final ', <one word>'
    becomes ' <andword> <word>'; whitespace collapsed."""
    if num == 0:
        return "zero"
    digits = str(num)
    groups = []
    while digits:
        groups.append(int(digits[-3:]))
        digits = digits[:-3]
    if len(groups) > len(_MILL):
        raise ValueError("number too large to spell: %d" % num)
    parts = []
    for idx in range(len(groups) - 1, -1, -1):
        g = groups[idx]
        h, t, u = g // 100, (g // 10) % 10, g % 10
        if h:
            a = (" %s " % andword) if (t or u) else ""
            parts.append("%s hundred%s%s%s" % (_UNITS[h], a, _tenfn(t, u), _MILL[idx]))
        elif t or u:
            parts.append("%s%s" % (_tenfn(t, u), _MILL[idx]))
    out = ", ".join(parts)
    out = re.sub(r", (\S+)\s*\Z", " %s \\1" % andword, out)
    return _whitespace_re.sub(" ", out).strip()


def _pairs(num, zero):
    """inflect.number_to_words(num, andword='', zero=zero, group=2): digits read in pairs from the left."""
    digits = str(num)
    out = []
    i = 0
    while i < len(digits):
        if i + 1 < len(digits):
            t, u = int(digits[i]), int(digits[i + 1])
            if t:
                out.append(_tenfn(t, u))
            elif u:
                out.append("%s %s" % (zero, _UNITS[u]))
            else:
                out.append("%s %s" % (zero, zero))
            i += 2
        else:
            u = int(digits[i])
            out.append(_UNITS[u] if u else zero)
            i += 1
    return ", ".join(out)


def _ordinal_words(num):
    words = _cardinal(num, "and")
    if _ordinal_suffix_re.search(words):
        return _ordinal_suffix_re.sub(lambda m: _ORDINAL[m.group(1)], words)
    return words + "th"


_comma_number_re = re.compile(r"([0-9][0-9\,]+[0-9])")
_decimal_number_re = re.compile(r"([0-9]+\.[0-9]+)")
_pounds_re = re.compile(r"£([0-9\,]*[0-9]+)")
_dollars_re = re.compile(r"\$([0-9\.\,]*[0-9]+)")
_ordinal_re = re.compile(r"[0-9]+(st|nd|rd|th)")
_number_re = re.compile(r"[0-9]+")


def _expand_dollars(m):                     # tokenizer.py:58-78
    match = m.group(1)
    parts = match.split(".")
    if len(parts) > 2:
        return match + " dollars"
    dollars = int(parts[0]) if parts[0] else 0
    cents = int(parts[1]) if len(parts) > 1 and parts[1] else 0
    if dollars and cents:
        return "%s %s, %s %s" % (dollars, "dollar" if dollars == 1 else "dollars", cents, "cent" if cents == 1 else "cents")
    if dollars:
        return "%s %s" % (dollars, "dollar" if dollars == 1 else "dollars")
    if cents:
        return "%s %s" % (cents, "cent" if cents == 1 else "cents")
    return "zero dollars"


def _expand_number(m):                      # tokenizer.py:85-98
    num = int(m.group(0))
    if 1000 < num < 3000:
        if num == 2000:
            return "two thousand"
        if 2000 < num < 2010:
            return "two thousand " + _cardinal(num % 100, "and")
        if num % 100 == 0:
            return _cardinal(num // 100, "and") + " hundred"
        return _pairs(num, "oh").replace(", ", " ")
    return _cardinal(num, "")


def normalize_numbers(text):                # tokenizer.py:101-108
    text = re.sub(_comma_number_re, lambda m: m.group(1).replace(",", ""), text)
    text = re.sub(_pounds_re, r"\1 pounds", text)
    text = re.sub(_dollars_re, _expand_dollars, text)
    text = re.sub(_decimal_number_re, lambda m: m.group(1).replace(".", " point "), text)
    text = re.sub(_ordinal_re, lambda m: _ordinal_words(int(re.match(r"[0-9]+", m.group(0)).group(0))), text)
    text = re.sub(_number_re, _expand_number, text)
    return text


def expand_abbreviations(text):             # tokenizer.py:37-40
    for regex, replacement in _abbreviations:
        text = re.sub(regex, replacement, text)
    return text


_ASCII_TABLE = {
    "‘": "'", "’": "'", "‚": ",", "‛": "'", "“": '"', "”": '"', "„": '"', "′": "'",
    "″": '"', "‐": "-", "‑": "-", "‒": "-", "–": "-", "—": "--", "―": "--",
    "…": "...", " ": " ", " ": " ", " ": " ", " ": " ", "​": "", "«": "<<", "»": ">>",
    "ß": "ss", "æ": "ae", "Æ": "AE", "œ": "oe", "Œ": "OE", "ø": "o", "Ø": "O",
    "đ": "d", "Đ": "D", "ł": "l", "Ł": "L", "þ": "th", "Þ": "Th", "ð": "d", "Ð": "D",
    "€": "EUR", "¢": "C/", "¥": "Y=", "©": "(c)", "®": "(r)", "™": "(tm)", "°": "deg",
    "×": "x", "÷": "/", "•": "*", "·": "*", "½": " 1/2", "¼": " 1/4", "¾": " 3/4",
    "¡": "!", "¿": "?",
}


def to_ascii(text):
    if text.isascii():
        return text
    out = []
    for ch in text:
        if ord(ch) < 128 or ch == "£":       # '£' survives unidecode? no: unidecode('£') == 'PS'; see below
            out.append(ch)
            continue
        if ch in _ASCII_TABLE:
            out.append(_ASCII_TABLE[ch])
            continue
        dec = unicodedata.normalize("NFKD", ch)
        out.append("".join(c for c in dec if ord(c) < 128))
    return "".join(out).replace("£", "PS")    # unidecode maps the pound sign to 'PS' (so '£5' never reaches _pounds_re)


def basic_cleaners(text):
    return _whitespace_re.sub(" ", text.lower())


def english_cleaners(text):
    text = to_ascii(text)
    text = text.lower()
    text = normalize_numbers(text)
    text = expand_abbreviations(text)
    text = _whitespace_re.sub(" ", text)
    return text.replace('"', "")

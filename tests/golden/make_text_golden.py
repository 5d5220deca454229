#!/usr/bin/env python
"""Generates tests/golden/text_split_v1.json: outputs of the REFERENCE `split_and_recombine_text`
(/root/reference/tortoise/utils/text.py:4-72) on seeded synthetic texts. Run where /root/reference exists."""
import importlib.util
import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))
WORDS = ("the quick brown fox jumps over a lazy dog while seventeen extraordinarily long-winded "
         "parliamentarians deliberate uncharacteristically about internationalisation and tea").split()


def synth_text(rng, n_sent):
    out = []
    for _ in range(n_sent):
        n = rng.randint(1, 28)
        s = " ".join(rng.choice(WORDS) for _ in range(n)).capitalize()
        r = rng.random()
        if r < 0.15:
            s = '"' + s + rng.choice([".", "!", "?", "...", ""]) + '"'
        elif r < 0.2:
            s = "“" + s + ".”"
        else:
            s += rng.choice([".", ".", ".", "!", "?", "?!", "!!!", "...", ";", ",", ""])
        out.append(s)
        out.append(rng.choice([" ", " ", "  ", "\n", "\n\n", "\n\n\n", " \t "]))
    return "".join(out)


def cases():
    rng = random.Random(1234)
    cs = []
    for i in range(36):
        text = synth_text(rng, rng.randint(1, 24))
        d = rng.choice([20, 30, 60, 120, 200])
        m = d + rng.choice([10, 20, 50, 100])
        cs.append({"text": text, "desired_length": d, "max_length": m})
    cs += [{"text": t, "desired_length": 200, "max_length": 300} for t in ["", " ", ".", '"', 'a', '"a."', "a. b", "?!?", "x" * 700]]
    return cs


def main():
    spec = importlib.util.spec_from_file_location("ref_text", "/root/reference/tortoise/utils/text.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = []
    for c in cases():
        c = dict(c)
        c["chunks"] = ref.split_and_recombine_text(c["text"], c["desired_length"], c["max_length"])
        out.append(c)
    with open(os.path.join(HERE, "text_split_v1.json"), "w") as f:
        json.dump(out, f, indent=0, ensure_ascii=True)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()

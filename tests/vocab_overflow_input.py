"""Input with more than 325 000 distinct words: triggers the reference's order-dependent prune."""
import numpy as np


def overflow_text(seed=5, vocab=420000, uniform=1000000, hot=500000):
    rng = np.random.default_rng(seed)
    lens = rng.integers(4, 8, vocab)
    letters = rng.integers(97, 123, (vocab, 7)).astype(np.uint8)
    occ = np.concatenate([rng.integers(0, vocab, uniform), rng.integers(0, 3000, hot)])
    rng.shuffle(occ)
    return b"".join(letters[w, :lens[w]].tobytes() + b" " for w in occ)

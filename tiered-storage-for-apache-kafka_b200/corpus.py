"""Deterministic synthetic segments: R = uniform random bytes (what the reference's own tests
use, TransformsEndToEndTest.java:38-41), K = Kafka-like compressible text (Zipf-distributed words from a seeded
4096-word dictionary, libzstd-3 ratio ~4-5), Z = zeros.  Chunk-independent seeding so any chunk can be
regenerated on its own."""
import numpy as np

SEED_BASE = 0x5EED0001


def _dictionary(seed):
    rng = np.random.default_rng(seed ^ 0xD1C7)
    lens = rng.integers(3, 10, 4096)
    words = [bytes(rng.integers(97, 123, int(n), dtype=np.uint8)) + b" " for n in lens]
    return words


_DICT_CACHE = {}


def gen_chunk(kind, segment, chunk, size):
    """bytes of chunk `chunk` of segment `segment` (np.uint8 array of `size`)."""
    seed = (SEED_BASE + segment) * 1000003 + chunk
    if kind == "Z":
        return np.zeros(size, dtype=np.uint8)
    rng = np.random.default_rng(seed)
    if kind == "R":
        return rng.integers(0, 256, size, dtype=np.uint8)
    if kind == "K":
        key = SEED_BASE
        if key not in _DICT_CACHE:
            words = _dictionary(key)
            maxlen = max(len(w) for w in words)
            tab = np.zeros((4096, maxlen), dtype=np.uint8)
            wl = np.zeros(4096, dtype=np.int64)
            for i, w in enumerate(words):
                tab[i, :len(w)] = np.frombuffer(w, dtype=np.uint8)
                wl[i] = len(w)
            ranks = np.arange(1, 4097, dtype=np.float64)
            p = ranks ** -1.3
            p /= p.sum()
            _DICT_CACHE[key] = (tab, wl, np.cumsum(p))
        tab, wl, cdf = _DICT_CACHE[key]
        nwords = size // 4 + 16
        ids = np.searchsorted(cdf, rng.random(nwords)).clip(0, 4095)
        lens = wl[ids]
        ends = np.cumsum(lens)
        total = int(ends[-1])
        starts = ends - lens
        # scatter words: position j of the output belongs to word w(j) at offset j - starts[w(j)]
        widx = np.repeat(np.arange(nwords), lens)
        within = np.arange(total) - np.repeat(starts, lens)
        out = tab[ids[widx], within]
        assert total >= size
        return np.ascontiguousarray(out[:size])
    raise ValueError(kind)


def gen_segment(kind, segment, size, chunk_size):
    out = np.empty(size, dtype=np.uint8)
    n = (size + chunk_size - 1) // chunk_size if size else 0
    for i in range(n):
        lo = i * chunk_size
        hi = min(size, lo + chunk_size)
        out[lo:hi] = gen_chunk(kind, segment, i, hi - lo)
    return out


def fixed_key_material(n_chunks):
    """C3 of SURVEY.md §8d: key = 00..1f, AAD = 32 x A5, IV_i = BE96(i)."""
    key = bytes(range(32))
    aad = bytes([0xA5]) * 32
    ivs = b"".join(int(i).to_bytes(12, "big") for i in range(n_chunks))
    return key, aad, ivs

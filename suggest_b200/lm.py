"""Host-side mirror of pkg/lm and pkg/spellchecker over the C ABI (sg_lm_*, sg_predict_batch).

The model itself lives in HBM and every score is computed by the CUDA kernels of csrc/sg_lm.cu; what stays on the host is
what the reference also keeps around the model: the vocabulary (word -> id), reading the Google n-gram text files into the
packed arrays (pkg/lm/ngram_reader.go, ngram_vector_builder.go, packed_array.go:198-237) and splitting a sentence into
words (pkg/lm/tokenizer.go, pkg/analysis/word_tokenizer.go).
"""
import ctypes as C
from typing import List, Sequence

import numpy as np

from . import _capi

UnknownWordID = 0xFFFFFFFF          # pkg/lm/indexer.go:18
UnknownWordScore = -100.0           # pkg/lm/ngram_model.go:24
InvalidContextOffset = 0xFFFFFFFD   # pkg/lm/ngram_vector.go:31-35


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _pack_ids(rows):
    off = np.zeros(len(rows) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(r) for r in rows])
    flat = np.fromiter((x for r in rows for x in r), dtype=np.uint32, count=int(off[-1]))
    return flat, off


class Indexer:
    """lm.Indexer (pkg/lm/indexer.go:22-29): two-way mapping between words and ids.  The reference answers Get through a
    minimal perfect hash plus a dictionary probe; a dict gives the same answers."""

    def __init__(self, words: Sequence[str]):
        self.words = list(words)
        self._ids = {w: i for i, w in enumerate(self.words)}

    def Get(self, token: str) -> int:
        return self._ids.get(token, UnknownWordID)

    def Find(self, index: int) -> str:
        return self.words[index] if 0 <= index < len(self.words) else "<UNK>"


def read_unigram_words(text: str, binary_order=False) -> List[str]:
    """vocabulary of a "1-gm" file: line order (indexer.go:85-113) or (count desc, word asc) as lm's binary build numbers
    it (binary.go:136-189)"""
    items = []
    for line in text.split("\n"):
        if not line:
            continue
        tab = line.index("\t")
        if binary_order and tab == 0:
            continue
        items.append((line[:tab], int(line[tab + 1:])))
    if binary_order:
        items = sorted(set(items), key=lambda wc: (-wc[1], wc[0].encode()))
    return [w for w, _ in items]


def levels_from_sentences(sentences: Sequence[Sequence[int]], order: int, start: int, end: int):
    """The packed arrays of every level from word-id sentences: what lm's build path produces from a corpus — every
    sentence wrapped in <S> ... </S>, all k-grams for k = 1..order counted (NGramBuilder.Build, pkg/lm/ngram_builder.go:21-42),
    then filed under (context offset, word) like NGramVectorBuilder + CreatePackedArray (ngram_vector_builder.go:69-106,
    packed_array.go:198-237).  Vectorised with numpy sorts instead of a count trie and red-black trees."""
    lens = np.fromiter((len(s) + 2 for s in sentences), dtype=np.int64, count=len(sentences))
    flat = np.empty(int(lens.sum()), dtype=np.uint64)
    starts = np.zeros(len(sentences) + 1, dtype=np.int64)
    starts[1:] = np.cumsum(lens)
    flat[starts[:-1]] = start
    flat[starts[1:] - 1] = end
    body = np.ones(len(flat), dtype=bool)
    body[starts[:-1]] = False
    body[starts[1:] - 1] = False
    flat[body] = np.fromiter((w for s in sentences for w in s), dtype=np.uint64, count=int(body.sum()))
    sent_of = np.repeat(np.arange(len(sentences)), lens)
    levels, prev_keys = [], None
    ctx_of_pos = np.full(len(flat), InvalidContextOffset, dtype=np.uint64)  # context offset of the (k-1)-gram starting at pos
    for k in range(1, order + 1):
        n = len(flat) - k + 1
        if n <= 0:
            levels.append((np.zeros(0, np.uint64), np.zeros(0, np.uint64), 0))
            continue
        ok = sent_of[:n] == sent_of[k - 1:k - 1 + n]           # the k-gram stays inside one sentence
        pos = np.flatnonzero(ok)
        keys = ctx_of_pos[pos] << np.uint64(32) | flat[pos + k - 1]
        uniq, inverse, counts = np.unique(keys, return_inverse=True, return_counts=True)
        ctx = uniq >> np.uint64(32)
        values = (uniq & np.uint64(0xFFFFFFFF)) << np.uint64(32) | (counts.astype(np.uint64) & np.uint64(0xFFFFFFFF))
        first = np.ones(len(uniq), dtype=bool)
        first[1:] = ctx[1:] != ctx[:-1]
        containers = ctx[first] << np.uint64(32) | np.flatnonzero(first).astype(np.uint64)
        levels.append((containers, values, int(counts.sum()) & 0xFFFFFFFF))
        ctx_of_pos = np.full(len(flat), InvalidContextOffset, dtype=np.uint64)
        ctx_of_pos[pos] = inverse.astype(np.uint64)            # offset of this k-gram in its level = context of the (k+1)-gram
    return levels


class NGramModel:
    """lm.NGramModel on the device."""

    def __init__(self, handle, order):
        self._h, self.order = handle, order

    @classmethod
    def from_levels(cls, levels, device=0):
        """levels[i] = (containers uint64[], values uint64[], total) of the (i+1)-grams"""
        order = len(levels)
        keep = [(np.ascontiguousarray(c, dtype=np.uint64), np.ascontiguousarray(v, dtype=np.uint64)) for c, v, _ in levels]
        cp = (C.c_void_p * order)(*[c.ctypes.data for c, _ in keep])
        vp = (C.c_void_p * order)(*[v.ctypes.data for _, v in keep])
        nc = np.array([len(c) for c, _ in keep], dtype=np.uint64)
        nv = np.array([len(v) for _, v in keep], dtype=np.uint64)
        totals = np.array([t & 0xFFFFFFFF for _, _, t in levels], dtype=np.uint32)
        h = C.c_void_p()
        _capi.check(_capi.lib().sg_lm_create(order, cp, _ptr(nc), vp, _ptr(nv), _ptr(totals), device, C.byref(h)))
        return cls(h, order)

    @classmethod
    def from_google_ngrams(cls, files: Sequence[str], indexer: Indexer, device=0):
        """NewGoogleNGramReader(order, indexer, directory).Read() (pkg/lm/ngram_reader.go:37-98) from the texts of 1-gm ... n-gm"""
        levels, prev = [], None
        for order, text in enumerate(files, start=1):
            nodes = {}
            for line in text.split("\n"):
                if not line:
                    continue
                tab = line.index("\t")
                ids = [indexer.Get(w) for w in line[:tab].split(" ")]
                if len(ids) != order:
                    raise ValueError("nGrams order is out of range")
                parent = InvalidContextOffset
                for lvl, w in enumerate(ids[:-1]):  # context offset = position of the prefix in the level below
                    parent = levels_lookup[lvl].get((parent, w), InvalidContextOffset)  # unigrams live under InvalidContextOffset
                key = parent << 32 | ids[-1]
                nodes[key] = (nodes.get(key, 0) + int(line[tab + 1:])) & 0xFFFFFFFF
            keys = np.array(sorted(nodes), dtype=np.uint64)
            counts = np.array([nodes[int(k_)] for k_ in keys], dtype=np.uint64)
            values = (keys & np.uint64(0xFFFFFFFF)) << np.uint64(32) | counts
            ctx = keys >> np.uint64(32)
            first = np.ones(len(keys), dtype=bool)
            first[1:] = ctx[1:] != ctx[:-1]
            containers = ctx[first] << np.uint64(32) | np.flatnonzero(first).astype(np.uint64)
            total = int(counts.sum()) & 0xFFFFFFFF
            levels.append((containers, values, total))
            lookup = {(int(c), int(k_ & np.uint64(0xFFFFFFFF))): i for i, (c, k_) in enumerate(zip(ctx, keys))}
            if order == 1:
                levels_lookup = [lookup]
            else:
                levels_lookup.append(lookup)
        return cls.from_levels(levels, device)

    @classmethod
    def open(cls, path, device=0):
        """nGramModel.Load of a binary model file (pkg/lm/ngram_model.go:126-160)"""
        h = C.c_void_p()
        _capi.check(_capi.lib().sg_lm_open(str(path).encode(), device, C.byref(h)))
        with open(path, "rb") as f:
            order = f.read(6)[5]
        return cls(h, order)

    def close(self):
        if self._h is not None:
            _capi.lib().sg_lm_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if self._h is None:
            raise _capi.SuggestError(_capi.SG_ERR_INVALID, "model is closed")
        return self._h

    def ScoreBatch(self, ngrams: Sequence[Sequence[int]]) -> np.ndarray:
        """nGramModel.Score for every n-gram (pkg/lm/ngram_model.go:44-64)"""
        ids, off = _pack_ids(ngrams)
        out = np.zeros(len(ngrams), dtype=np.float64)
        _capi.check(_capi.lib().sg_lm_score_batch(self.handle, _ptr(ids), _ptr(off), len(ngrams), _ptr(out)))
        return out

    def Score(self, ngram: Sequence[int]) -> float:
        return float(self.ScoreBatch([ngram])[0])

    def ScoreNextBatch(self, contexts: Sequence[Sequence[int]], candidates: Sequence[Sequence[int]]):
        """Next(context).ScoreNext(candidate) for every candidate of every query -> (list of score arrays, has_scorer)"""
        ctx, ctx_off = _pack_ids(contexts)
        cand, cand_off = _pack_ids(candidates)
        out = np.zeros(len(cand), dtype=np.float64)
        has = np.zeros(len(contexts), dtype=np.uint8)
        _capi.check(_capi.lib().sg_lm_score_next_batch(self.handle, _ptr(ctx), _ptr(ctx_off), len(contexts), _ptr(cand), _ptr(cand_off),
                                                       _ptr(out), _ptr(has)))
        return [out[cand_off[i]:cand_off[i + 1]] for i in range(len(contexts))], has.astype(bool)


class LanguageModel:
    """lm.LanguageModel (pkg/lm/language_model.go)"""

    def __init__(self, model: NGramModel, indexer: Indexer, ngram_order: int, start_symbol="<S>", end_symbol="</S>"):
        self.model, self.indexer, self.order = model, indexer, ngram_order
        self.start, self.end = indexer.Get(start_symbol), indexer.Get(end_symbol)
        if UnknownWordID in (self.start, self.end):
            raise ValueError("failed to get wordID of the start / end symbol")

    def GetWordID(self, token):
        return self.indexer.Get(token)

    def ScoreWordIDs(self, sequence):
        seq = [self.start] + list(sequence) + [self.end]
        k = self.order
        grams = [seq[i:i + k] for i in range(len(seq) - k + 1)] if len(seq) >= k else []  # splitIntoNGrams, generator.go:9-23
        return float(self.model.ScoreBatch(grams).sum()) if grams else 0.0

    def ScoreSentence(self, sentence):
        return self.ScoreWordIDs([self.indexer.Get(t) for t in sentence])

    def next_context(self, sequence):
        """the sequence languageModel.Next hands to nGramModel.Next (language_model.go:103-115)"""
        seq, k = list(sequence), self.order
        if len(seq) + 1 < k:
            seq = [self.start] + seq
        elif len(seq) > k:
            seq = seq[len(seq) - k + 1:]
        elif len(seq) == k:
            seq = seq[:k - 1]
        return seq


def word_tokenize(text: str, has) -> List[str]:
    """lm.NewTokenizer(alphabet).Tokenize (pkg/lm/tokenizer.go:24-31, pkg/analysis/word_tokenizer.go:22-48)"""
    text = text.lower().strip(" ")
    words, cur = [], ""
    for ch in text:
        if has(ch):
            cur += ch
        else:
            if cur:
                words.append(cur)
            cur = ""
    if cur:
        words.append(cur)
    return words


class SpellChecker:
    """spellchecker.SpellChecker (pkg/spellchecker/spellchecker.go:17-38): index over the vocabulary, language model,
    sentence tokenizer, dictionary (here the indexer's word list)."""

    def __init__(self, index, model: LanguageModel, tokenizer, words: Sequence[str]):
        self.index, self.model, self.tokenizer, self.words = index, model, tokenizer, list(words)

    def PredictBatch(self, queries: Sequence[str], topK: int, similarity: float) -> List[List[str]]:
        """Predict for every query through sg_predict_batch"""
        from .suggest import pack_strings
        rows, last, ctxs = [], [], []
        for q in queries:
            tokens = self.tokenizer(q)
            if not tokens:
                rows.append(None)
                continue
            rows.append(len(last))
            last.append(tokens[-1])
            seq = [self.model.GetWordID(t) for t in tokens[:-1]]
            ctxs.append(self.model.next_context(seq) if seq else [])  # no context: no scorer (spellchecker.go:102-104)
        out = [[] for _ in queries]
        if last:
            if topK <= 0:
                raise _capi.SuggestError(_capi.SG_ERR_INVALID, "topK is invalid")
            data, off = pack_strings(last)
            ctx, ctx_off = _pack_ids(ctxs)
            n, k = len(last), int(topK)
            ids = np.zeros((n, k + 1), dtype=np.uint32)
            cnt = np.zeros(n, dtype=np.uint32)
            _capi.check(_capi.lib().sg_predict_batch(self.index.handle, self.model.model.handle, _ptr(data), _ptr(off.astype(np.uint32)),
                                                     _ptr(ctx), _ptr(ctx_off), n, float(similarity), k, _ptr(ids), _ptr(cnt)))
            for i, r in enumerate(rows):
                if r is not None:
                    out[i] = [self.words[int(d)] for d in ids[r, :int(cnt[r])]]
        return out

    def Predict(self, query: str, topK: int, similarity: float) -> List[str]:
        return self.PredictBatch([query], topK, similarity)[0]

    def predict_ids(self, last_words, contexts, topK, similarity):
        """sg_predict_batch on already tokenised input -> list of id lists (used by the parity tests)"""
        from .suggest import pack_strings
        data, off = pack_strings(last_words)
        ctx, ctx_off = _pack_ids(contexts)
        n, k = len(last_words), int(topK)
        ids = np.zeros((n, k + 1), dtype=np.uint32)
        cnt = np.zeros(n, dtype=np.uint32)
        _capi.check(_capi.lib().sg_predict_batch(self.index.handle, self.model.model.handle, _ptr(data), _ptr(off.astype(np.uint32)),
                                                 _ptr(ctx), _ptr(ctx_off), n, float(similarity), k, _ptr(ids), _ptr(cnt)))
        return [[int(d) for d in ids[i, :int(cnt[i])]] for i in range(n)]

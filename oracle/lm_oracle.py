"""CPU restatement of the reference's n-gram language model and spellchecker.  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import it; the product path never does.

Follows, function by function:
  pkg/lm/ngram_vector_builder.go:40-106  Key = context << 32 | word, nodes ordered by key, parents resolved per prefix
  pkg/lm/packed_array.go:198-237         CreatePackedArray: values[] = word << 32 | count, containers[] = context << 32 | from
  pkg/lm/packed_array.go:52-60,163-210   GetCount / find / findContainerPos (two binary searches)
  pkg/lm/ngram_model.go:44-99,163-175    Score (stupid backoff, alpha = 0.4), Next, calcScore
  pkg/lm/scorer_next.go:15-23            ScoreNext
  pkg/lm/language_model.go:72-132        ScoreWordIDs / Next (sentence wrapping), generator.go:9-23 splitIntoNGrams
  pkg/lm/ngram_reader.go:37-98           Google n-gram text format "w1 w2 w3\\tcount"
  pkg/lm/ngram_model.go:101-160, packed_array.go:98-160   binary ".lm" format
  pkg/lm/binary.go:136-189               vocabulary order of the binary build: (count desc, word asc)
  pkg/analysis/word_tokenizer.go:22-48, pkg/lm/tokenizer.go:24-31   sentence -> words
  pkg/spellchecker/spellchecker.go:40-151, collector.go:61-78, scorer.go:17-41   Predict
Parity status: pinned to pkg/lm/ngram_model_test.go:28-149 and language_model_test.go:12-70 (tests/test_lm_oracle.py),
including the shipped binary fixture test.lm; pkg/spellchecker has no tests in the reference ("TODO add tests!!").
"""
import bisect
import math
import struct

INVALID_CONTEXT = 0xFFFFFFFD  # maxUint32 - 2, pkg/lm/ngram_vector.go:31-35
UNKNOWN_WORD_ID = 0xFFFFFFFF
UNKNOWN_WORD_SCORE = -100.0
ALPHA = 0.4


class PackedArray:
    """pkg/lm/packed_array.go"""

    def __init__(self, containers, values, total):
        self.containers, self.values, self.total = list(containers), list(values), int(total)

    @classmethod
    def from_nodes(cls, nodes):  # nodes: iterable of (key, count) in key order
        containers, values, total, context = [], [], 0, INVALID_CONTEXT
        for frm, (key, count) in enumerate(nodes):
            total = (total + count) & 0xFFFFFFFF
            cur = key >> 32
            if context != cur or not containers:
                containers.append(cur << 32 | frm)
                context = cur
            values.append((key & 0xFFFFFFFF) << 32 | count)
        return cls(containers, values, total)

    def container_pos(self, context):
        c = self.containers
        if not c or (c[0] >> 32) > context or (c[-1] >> 32) < context:
            return -1
        i = bisect.bisect_left(c, context << 32)
        if i >= len(c) or (c[i] >> 32) != context:
            return -1
        return i

    def find(self, word, context):
        i = self.container_pos(context)
        if i == -1:
            return 0, INVALID_CONTEXT
        frm = self.containers[i] & 0xFFFFFFFF
        to = len(self.values) if i == len(self.containers) - 1 else self.containers[i + 1] & 0xFFFFFFFF
        vals = self.values[frm:to]
        if (vals[0] >> 32) > word or (vals[-1] >> 32) < word:
            return 0, INVALID_CONTEXT
        j = bisect.bisect_left(vals, word << 32)
        if j >= len(vals) or (vals[j] >> 32) != word:
            return 0, INVALID_CONTEXT
        return vals[j], frm + j

    def get_count(self, word, context):
        v, off = self.find(word, context)
        return (0, INVALID_CONTEXT) if off == INVALID_CONTEXT else (v & 0xFFFFFFFF, off)

    def sub_range(self, context):
        """SubVector(context): [from, to) of the values of that context, or None"""
        i = self.container_pos(context)
        if i == -1:
            return None
        frm = self.containers[i] & 0xFFFFFFFF
        to = len(self.values) if i == len(self.containers) - 1 else self.containers[i + 1] & 0xFFFFFFFF
        return frm, to


def calc_score(counts):
    factor = 1.0
    for i in range(len(counts) - 1, 0, -1):
        if counts[i] > 0:
            return math.log(factor * float(counts[i]) / float(counts[i - 1]))
        factor *= ALPHA
    return UNKNOWN_WORD_SCORE


class NGramModel:
    """pkg/lm/ngram_model.go"""

    def __init__(self, vectors):
        self.vectors = vectors
        self.order = len(vectors)

    def score(self, ngrams):
        order = min(self.order, len(ngrams))
        counts = [0] * (order + 1)
        parent = INVALID_CONTEXT
        for i in range(order):
            if i == 0:
                counts[0] = self.vectors[0].total
            counts[i + 1], parent = self.vectors[i].get_count(ngrams[i], parent)
        return calc_score(counts)

    def next(self, ngrams):
        """-> (context counts, context offset, level) or None; raises like the reference on a bad length"""
        if self.order <= len(ngrams) or len(ngrams) == 0:
            raise ValueError("nGrams length should be less than the nGramModel order")
        counts, parent = [], INVALID_CONTEXT
        for order, w in enumerate(ngrams):
            count, parent = self.vectors[order].get_count(w, parent)
            if count == 0:
                return None
            counts.append(count)
        level = len(ngrams)
        if self.vectors[level].sub_range(parent) is None:
            return None
        return counts, parent, level

    def score_next(self, nxt, word):
        """scorerNext.ScoreNext"""
        counts, parent, level = nxt
        count, _ = self.vectors[level].get_count(word, parent)
        if count == 0:
            return UNKNOWN_WORD_SCORE
        return calc_score(counts + [count])

    # binary form
    def store(self):
        out = b"0.0.2" + bytes([self.order])
        for v in self.vectors:
            out += f"{8 * len(v.containers)} {8 * len(v.values)} {v.total}\n".encode()
            out += struct.pack(f"<{len(v.containers)}Q", *v.containers) + struct.pack(f"<{len(v.values)}Q", *v.values)
        return out

    @classmethod
    def load(cls, data):
        if data[:5] != b"0.0.2":
            raise ValueError("Version mismatch")
        order, p, vectors = data[5], 6, []
        for _ in range(order):
            nl = data.index(b"\n", p)
            csize, vsize, total = (int(x) for x in data[p:nl].split())
            p = nl + 1
            containers = struct.unpack(f"<{csize // 8}Q", data[p:p + csize])
            values = struct.unpack(f"<{vsize // 8}Q", data[p + csize:p + csize + vsize])
            p += csize + vsize
            vectors.append(PackedArray(containers, values, total))
        return cls(vectors)


def read_google_ngrams(files, word_id):
    """files: text of "<order>-gm" for order 1..n; word_id: token -> id (UNKNOWN_WORD_ID if absent)."""
    vectors = []
    for order, text in enumerate(files, start=1):
        tree = {}
        for line in text.split("\n"):
            if not line:
                continue
            tab = line.index("\t")
            ids = [word_id(w) for w in line[:tab].split(" ")]
            count = int(line[tab + 1:])
            if len(ids) != order:
                raise ValueError("nGrams order is out of range")
            parent = INVALID_CONTEXT
            for i, w in enumerate(ids[:-1]):
                _, parent = vectors[i].find(w, parent)
            key = parent << 32 | ids[-1]
            tree[key] = (tree.get(key, 0) + count) & 0xFFFFFFFF
        vectors.append(PackedArray.from_nodes(sorted(tree.items())))
    return NGramModel(vectors)


def vocabulary_from_unigrams(text, binary_order=False):
    """word list whose index is the word id: line order (buildIndexerWithInMemoryDictionary, indexer.go:85-113) or the
    (count desc, word asc) order of the binary build (binary.go:136-189)"""
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


class LanguageModel:
    """pkg/lm/language_model.go"""

    def __init__(self, model, words, order, start="<S>", end="</S>"):
        self.model, self.words, self.order = model, list(words), order
        self.ids = {w: i for i, w in enumerate(self.words)}
        self.start, self.end = self.word_id(start), self.word_id(end)

    def word_id(self, token):
        return self.ids.get(token, UNKNOWN_WORD_ID)

    def score_word_ids(self, seq):
        seq = [self.start] + list(seq) + [self.end]
        k = self.order
        if len(seq) < k:
            return 0.0
        return sum(self.model.score(seq[i:i + k]) for i in range(len(seq) - k + 1))

    def score_sentence(self, sentence):
        return self.score_word_ids([self.word_id(t) for t in sentence])

    def next_context(self, seq):
        """the sequence languageModel.Next hands to nGramModel.Next (language_model.go:103-115)"""
        seq = list(seq)
        k = self.order
        if len(seq) + 1 < k:
            seq = [self.start] + seq
        elif len(seq) > k:
            seq = seq[len(seq) - k + 1:]
        elif len(seq) == k:
            seq = seq[:k - 1]
        return seq

    def next(self, seq):
        return self.model.next(self.next_context(seq))


def word_tokenize(text, has):
    """lm.NewTokenizer: strings.ToLower, Trim(" "), then runs of alphabet members (word_tokenizer.go:22-48).
    `text` is a str; lowering is Python's, which equals Go's for the test alphabets."""
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


def predict(index_oracle, lm, tokens, top_k, similarity, cosine_code, canonical_mode):
    """spellchecker.Predict (spellchecker.go:40-92) on already tokenised input -> list of candidate ids.
    index_oracle: oracle.OracleIndex over the LM vocabulary (document id = word id)."""
    if not tokens:
        return []
    word, seq = tokens[-1], tokens[:-1]
    seq_ids = [lm.word_id(t) for t in seq]
    nxt = lm.next(seq_ids) if seq_ids else None

    def score(doc):
        return lm.model.score_next(nxt, doc) if nxt is not None else UNKNOWN_WORD_SCORE

    # Autocomplete with the lm collector: every document holding all prefix n-grams, top-k by (score desc, id asc)
    all_ids, _ = index_oracle.autocomplete(word, max(index_oracle.n_docs, 1))
    ranked = sorted(((-score(int(d)), int(d)) for d in all_ids))[:top_k]
    candidates = [d for _, d in ranked]
    if len(candidates) < top_k:
        ids, _, n = index_oracle.suggest_batch([word], cosine_code, similarity, top_k, canonical_mode, threads=1)
        for d in ids[0, :int(n[0])]:
            if int(d) not in candidates:
                candidates.append(int(d))
    if nxt is not None:
        candidates = sorted(candidates, key=lambda d: -score(d))  # sort.SliceStable, score desc
    if top_k < len(candidates):
        candidates = candidates[:top_k + 1]  # sic, spellchecker.go:87-89
    return candidates

"""pkg/lm and pkg/spellchecker on the device (sg_lm_*, sg_predict_batch) against the reference's expectations and the
CPU oracle.  Needs a B200: `pytest -m gpu`."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import lm_oracle as LM
from oracle import oracle as O
import suggest_b200 as S
from suggest_b200 import lm as P
from suggest_b200.suggest import IndexDescription

pytestmark = pytest.mark.gpu
FIX = os.path.join(GOLDEN, "lm")
TOL = 1e-4          # the reference's own tolerance (pkg/lm/ngram_model_test.go:12)
TIGHT = 1e-12       # device log() vs libm log(), relative


def read(name):
    with open(os.path.join(FIX, name)) as f:
        return f.read()


def close(a, b):
    return abs(a - b) <= TIGHT * max(1.0, abs(a), abs(b))


@pytest.fixture(scope="module")
def sam():
    words = P.read_unigram_words(read("1-gm"))
    ix = P.Indexer(words)
    model = P.NGramModel.from_google_ngrams([read("1-gm"), read("2-gm"), read("3-gm")], ix)
    return model, ix


def test_score_kats(sam):  # pkg/lm/ngram_model_test.go:121-149
    model, ix = sam
    kat = [(["i", "am", "sam"], -0.6931), (["i", "am"], -0.4054), (["sam", "i", "am"], 0.0), (["sam", "am", "i"], -4.1351),
           (["i", "dont", "know"], -3.7297), (["no", "one", "word"], -100.0)]
    got = model.ScoreBatch([[ix.Get(w) for w in s] for s, _ in kat])
    for (s, want), g in zip(kat, got):
        assert abs(g - want) < TOL, (s, g, want)


def test_score_next_kats(sam):  # pkg/lm/ngram_model_test.go:28-88
    model, ix = sam
    kat = [(["i", "am"], "sam", -0.6931), (["i", "am"], "</S>", -0.6931), (["i"], "am", -0.4054), (["i"], "do", -1.0986),
           (["green"], "eggs", 0.0)]
    scores, has = model.ScoreNextBatch([[ix.Get(w) for w in c] for c, _, _ in kat], [[ix.Get(w)] for _, w, _ in kat])
    assert has.all()
    for (c, w, want), g in zip(kat, scores):
        assert abs(g[0] - want) < TOL, (c, w, g)
    scores, has = model.ScoreNextBatch([[ix.Get("i"), ix.Get("am")], [ix.Get("ham"), ix.Get("i")], [], [0, 1, 2]],
                                       [[ix.Get("ham"), ix.Get("sam")], [ix.Get("am")], [ix.Get("am")], [ix.Get("am")]])
    assert list(has) == [True, False, False, False]
    assert scores[0][0] == P.UnknownWordScore and abs(scores[0][1] + 0.6931) < TOL
    assert scores[1][0] == scores[2][0] == scores[3][0] == P.UnknownWordScore


def test_binary_model_and_sentences():  # language_model_test.go:31-70 (RetrieveLMFromBinary path)
    model = P.NGramModel.open(os.path.join(FIX, "test.lm"))
    assert model.order == 3
    ix = P.Indexer(P.read_unigram_words(read("1-gm"), binary_order=True))
    lm = P.LanguageModel(model, ix, 3)
    for sent, want in [(["i", "am", "sam"], -1.3862), (["i", "am"], -1.3862), (["sam", "i", "am"], -0.6931),
                       (["sam", "am", "i"], -10.2852), (["i", "dont", "know"], -105.0514), (["no", "one", "word"], -203.7297)]:
        assert abs(lm.ScoreSentence(sent) - want) < TOL, sent
    with pytest.raises(S.SuggestError):
        P.NGramModel.open(os.path.join(FIX, "1-gm"))


def synthetic_corpus(n_words, n_sentences, seed):
    rng = np.random.default_rng(seed)
    stems = ["".join(chr(97 + c) for c in rng.integers(0, 26, size=int(rng.integers(2, 5)))) for _ in range(max(8, n_words // 6))]
    words = sorted({stems[int(rng.integers(len(stems)))] + "".join(chr(97 + c) for c in rng.integers(0, 26, size=int(rng.integers(1, 6))))
                    for _ in range(n_words * 2)})[:n_words]
    zipf = np.minimum(rng.zipf(1.3, size=(n_sentences, 12)) - 1, len(words) - 1)
    lens = rng.integers(2, 12, size=n_sentences)
    sents = [[words[int(j)] for j in zipf[i, :lens[i]]] for i in range(n_sentences)]
    return words, sents


def google_files(words, sents, order=3):
    vocab = ["<S>", "</S>"] + words
    counts = [dict() for _ in range(order)]
    for s in sents:
        seq = ["<S>"] + s + ["</S>"]
        for k in range(1, order + 1):
            for i in range(len(seq) - k + 1):
                g = " ".join(seq[i:i + k])
                counts[k - 1][g] = counts[k - 1].get(g, 0) + 1
    for w in vocab:
        counts[0].setdefault(w, 1)
    files = ["".join(f"{g}\t{c}\n" for g, c in lvl.items()) for lvl in counts]
    return vocab, files


@pytest.fixture(scope="module")
def synthetic_lm():
    words, sents = synthetic_corpus(4000, 20000, 5)
    vocab, files = google_files(words, sents)
    vocab = P.read_unigram_words(files[0])
    ix = P.Indexer(vocab)
    model = P.NGramModel.from_google_ngrams(files, ix)
    ids = {w: i for i, w in enumerate(vocab)}
    omodel = LM.read_google_ngrams(files, lambda w: ids.get(w, LM.UNKNOWN_WORD_ID))
    return model, omodel, vocab, sents


def test_score_and_score_next_match_oracle(synthetic_lm):
    model, omodel, vocab, sents = synthetic_lm
    rng = np.random.default_rng(9)
    ids = {w: i for i, w in enumerate(vocab)}
    grams = []
    for s in sents[:3000]:
        seq = [ids["<S>"]] + [ids[w] for w in s] + [ids["</S>"]]
        i = int(rng.integers(0, len(seq)))
        g = seq[i:i + int(rng.integers(1, 5))]
        if rng.random() < 0.3:
            g[int(rng.integers(len(g)))] = int(rng.integers(len(vocab)))
        if rng.random() < 0.05:
            g[0] = P.UnknownWordID
        grams.append(g)
    got = model.ScoreBatch(grams)
    for g, v in zip(grams, got):
        assert close(v, omodel.score(g)), (g, v, omodel.score(g))
    ctxs, cands = [], []
    for s in sents[3000:5000]:
        seq = [ids["<S>"]] + [ids[w] for w in s]
        i = int(rng.integers(0, len(seq)))
        ctxs.append(seq[i:i + int(rng.integers(1, 3))])
        cands.append([int(x) for x in rng.integers(0, len(vocab), size=int(rng.integers(0, 20)))] + [ids[w] for w in s[:3]])
    scores, has = model.ScoreNextBatch(ctxs, cands)
    for c, cd, sc, h in zip(ctxs, cands, scores, has):
        nxt = omodel.next(c)
        assert h == (nxt is not None)
        for w, v in zip(cd, sc):
            want = omodel.score_next(nxt, w) if nxt is not None else LM.UNKNOWN_WORD_SCORE
            assert close(v, want), (c, w, v, want)


def test_predict_matches_oracle(synthetic_lm):
    """SpellChecker.Predict: completions ranked by the model, fuzzy top-up, stable re-sort, k + 1 cut (spellchecker.go:40-92)"""
    model, omodel, vocab, sents = synthetic_lm
    desc = dict(ngram_size=3, wrap=("$", "$"), pad="$", alphabet=("english", "$"))
    index = S.NewRAMBuilder(vocab, IndexDescription(Name="v", NGramSize=3, Alphabet=desc["alphabet"], Pad="$", Wrap=("$", "$"))).Build()
    ox = O.OracleIndex(3, desc["wrap"], desc["pad"], desc["alphabet"]).add_docs(vocab)
    ix = P.Indexer(vocab)
    lm = P.LanguageModel(model, ix, 3)
    olm = LM.LanguageModel(omodel, vocab, 3)
    sc = P.SpellChecker(index, lm, lambda t: t.split(" "), vocab)
    rng = np.random.default_rng(21)
    queries = []
    for s in sents[6000:6600]:
        cut = int(rng.integers(1, len(s) + 1))
        toks = list(s[:cut])
        w = toks[-1]
        mode = rng.random()
        if mode < 0.5:
            w = w[:max(1, int(rng.integers(1, len(w) + 1)))]               # a prefix: completions
        elif mode < 0.8:
            p_ = int(rng.integers(len(w)))
            w = w[:p_] + chr(97 + int(rng.integers(26))) + w[p_ + 1:]      # a typo: fuzzy candidates
        toks[-1] = w
        if rng.random() < 0.1:
            toks = toks[-1:]                                               # no context: no scorer
        if rng.random() < 0.05 and len(toks) > 1:
            toks[0] = "zzzzunknown"
        queries.append(toks)
    n_longer = 0
    for k, sim in ((5, 0.5), (1, 0.7), (12, 0.3), (3, 0.2)):
        last = [q[-1] for q in queries]
        ctxs = []
        for q in queries:
            seq = [lm.GetWordID(t) for t in q[:-1]]
            ctxs.append(lm.next_context(seq) if seq else [])
        got = sc.predict_ids(last, ctxs, k, sim)
        for q, g in zip(queries, got):
            want = LM.predict(ox, olm, q, k, sim, O.COSINE, O.CANONICAL)
            assert g == want, (q, k, sim, g, want)
            n_longer += len(g) > k
    assert n_longer > 0  # the k + 1 cut of spellchecker.go:87-89 is exercised
    words_out = sc.PredictBatch([" ".join(q) for q in queries[:50]] + ["", "   "], 5, 0.5)
    assert words_out[-1] == [] and words_out[-2] == []
    for q, row in zip(queries[:50], words_out):
        assert row == [vocab[d] for d in LM.predict(ox, olm, q, 5, 0.5, O.COSINE, O.CANONICAL)]
    index.close()


@pytest.mark.parametrize("order", [1, 2, 4])
def test_other_model_orders(order):
    """unigram, bigram and 4-gram models: every level of the context chain, contexts that are too long or empty"""
    words, sents = synthetic_corpus(300, 3000, 100 + order)
    ids = {w: i + 2 for i, w in enumerate(words)}
    id_sents = [[ids[w] for w in s] for s in sents]
    levels = P.levels_from_sentences(id_sents, order, 0, 1)
    model = P.NGramModel.from_levels(levels)
    omodel = LM.NGramModel([LM.PackedArray([int(x) for x in c], [int(x) for x in v], t) for c, v, t in levels])
    rng = np.random.default_rng(order)
    grams = [[int(x) for x in rng.integers(0, len(words) + 2, size=int(rng.integers(1, order + 3)))] for _ in range(300)]
    grams += [s[:order] for s in id_sents[:300]] + [[0] + s[:order - 1] for s in id_sents[:200] if order > 1]
    got = model.ScoreBatch(grams)
    for g, v in zip(grams, got):
        assert close(v, omodel.score(g)), (order, g, v, omodel.score(g))
    ctxs = [s[:int(rng.integers(0, order + 1))] for s in id_sents[300:700]]
    cands = [[int(x) for x in rng.integers(0, len(words) + 2, size=5)] + s[:4] for s in id_sents[300:700]]
    scores, has = model.ScoreNextBatch(ctxs, cands)
    for c, cd, sc, h in zip(ctxs, cands, scores, has):
        try:
            nxt = omodel.next(c)
        except ValueError:   # empty context, or as long as the model's order: the reference returns an error, no scorer here
            nxt = None
        assert h == (nxt is not None), (order, c)
        for w, v in zip(cd, sc):
            want = omodel.score_next(nxt, w) if nxt is not None else LM.UNKNOWN_WORD_SCORE
            assert close(v, want), (order, c, w, v, want)
    model.close()

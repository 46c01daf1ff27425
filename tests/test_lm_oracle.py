"""The LM / spellchecker oracle (oracle/lm_oracle.py) against the reference's own expectations.  No GPU.

pkg/lm/ngram_model_test.go:28-88 (ScoreNext), :121-149 (Score), :90-119 (binary round trip),
pkg/lm/language_model_test.go:50-70 (ScoreSentence, text and binary models).  Tolerance 1e-4 as in the reference."""
import math
import os

import pytest

from conftest import GOLDEN
from oracle import lm_oracle as LM

TOL = 1e-4
FIX = os.path.join(GOLDEN, "lm")


def read(name, mode="r"):
    with open(os.path.join(FIX, name), mode) as f:
        return f.read()


@pytest.fixture(scope="module")
def text_model():
    words = LM.vocabulary_from_unigrams(read("1-gm"))
    ids = {w: i for i, w in enumerate(words)}
    model = LM.read_google_ngrams([read("1-gm"), read("2-gm"), read("3-gm")], lambda w: ids.get(w, LM.UNKNOWN_WORD_ID))
    return model, words


SCORE_KAT = [(["i", "am", "sam"], -0.6931), (["i", "am"], -0.4054), (["sam", "i", "am"], 0.0), (["sam", "am", "i"], -4.1351),
             (["i", "dont", "know"], -3.7297), (["no", "one", "word"], -100.0)]


def check_model(model, words):
    ids = {w: i for i, w in enumerate(words)}
    for sent, want in SCORE_KAT:  # ngram_model_test.go:121-149
        got = model.score([ids.get(w, LM.UNKNOWN_WORD_ID) for w in sent])
        assert abs(got - want) < TOL, (sent, got, want)


def test_score_from_text(text_model):
    check_model(*text_model)


def test_score_next(text_model):  # ngram_model_test.go:28-88
    model, words = text_model
    ids = {w: i for i, w in enumerate(words)}
    for ctx, word, want in [(["i", "am"], "sam", -0.6931), (["i", "am"], "</S>", -0.6931), (["i"], "am", -0.4054),
                            (["i"], "do", -1.0986), (["green"], "eggs", 0.0)]:
        nxt = model.next([ids[w] for w in ctx])
        assert nxt is not None
        assert abs(model.score_next(nxt, ids[word]) - want) < TOL, (ctx, word)
    nxt = model.next([ids["i"], ids["am"]])
    assert model.score_next(nxt, ids["ham"]) == LM.UNKNOWN_WORD_SCORE      # unseen continuation: no back-off in ScoreNext
    assert model.next([ids["ham"], ids["i"]]) is None                        # unseen context
    with pytest.raises(ValueError):
        model.next([])
    with pytest.raises(ValueError):
        model.next([0, 1, 2])


def test_binary_round_trip_and_shipped_model(text_model):
    model, words = text_model
    again = LM.NGramModel.load(model.store())
    check_model(again, words)
    for a, b in zip(model.vectors, again.vectors):
        assert (a.containers, a.values, a.total) == (b.containers, b.values, b.total)
    # the reference's own binary build numbers the vocabulary by (count desc, word asc), binary.go:136-189
    shipped = LM.NGramModel.load(read("test.lm", "rb"))
    bwords = LM.vocabulary_from_unigrams(read("1-gm"), binary_order=True)
    assert bwords[:5] == ["</S>", "<S>", "i", "am", "sam"]
    check_model(shipped, bwords)
    ids = {w: i for i, w in enumerate(bwords)}
    rebuilt = LM.read_google_ngrams([read("1-gm"), read("2-gm"), read("3-gm")], lambda w: ids.get(w, LM.UNKNOWN_WORD_ID))
    blob = rebuilt.store()
    assert read("test.lm", "rb")[:len(blob)] == blob  # byte for byte; the file continues with the MPH table (binary.go:47-49)


SENTENCE_KAT = [(["i", "am", "sam"], -1.3862), (["i", "am"], -1.3862), (["sam", "i", "am"], -0.6931), (["sam", "am", "i"], -10.2852),
                (["i", "dont", "know"], -105.0514), (["no", "one", "word"], -203.7297)]


def test_score_sentence_text_and_binary(text_model):  # language_model_test.go:50-70
    model, words = text_model
    lm = LM.LanguageModel(model, words, 3)
    for sent, want in SENTENCE_KAT:
        assert abs(lm.score_sentence(sent) - want) < TOL, sent
    shipped = LM.LanguageModel(LM.NGramModel.load(read("test.lm", "rb")), LM.vocabulary_from_unigrams(read("1-gm"), True), 3)
    for sent, want in SENTENCE_KAT:
        assert abs(shipped.score_sentence(sent) - want) < TOL, sent


def test_next_context_wrapping(text_model):
    model, words = text_model
    lm = LM.LanguageModel(model, words, 3)
    i, am, sam = (lm.word_id(w) for w in ("i", "am", "sam"))
    assert lm.next_context([i]) == [lm.start, i]
    assert lm.next_context([i, am]) == [i, am]
    assert lm.next_context([sam, i, am]) == [sam, i]          # len == order: the last word is dropped (language_model.go:110-112)
    assert lm.next_context([sam, sam, i, am]) == [i, am]
    assert math.isclose(lm.model.score_next(lm.next([i]), am), math.log(1 / 2))  # <S> i am : <S> i = 1 : 2


def test_word_tokenizer():
    has = lambda ch: ch.isalnum() or ch in "-."  # noqa: E731
    assert LM.word_tokenize("  I am, Sam!  green-eggs ", has) == ["i", "am", "sam", "green-eggs"]
    assert LM.word_tokenize("", has) == []


def test_numpy_corpus_builder_equals_text_reader():
    """suggest_b200.lm.levels_from_sentences (host-side build, no GPU) = count the corpus, write Google n-gram text, read it
    back with the reference's reader rules"""
    import numpy as np
    from suggest_b200 import lm as P
    rng = np.random.default_rng(3)
    n_words = 50
    sents = [[int(w) + 2 for w in rng.integers(0, n_words, size=int(rng.integers(1, 9)))] for _ in range(400)]
    vocab = ["<S>", "</S>"] + [f"w{i}" for i in range(n_words)]
    levels = P.levels_from_sentences(sents, 3, 0, 1)
    counts = [dict() for _ in range(3)]
    for s in sents:
        seq = [0] + s + [1]
        for k in range(1, 4):
            for i in range(len(seq) - k + 1):
                g = " ".join(vocab[w] for w in seq[i:i + k])
                counts[k - 1][g] = counts[k - 1].get(g, 0) + 1
    files = ["".join(f"{g}\t{c}\n" for g, c in lvl.items()) for lvl in counts]
    ids = {w: i for i, w in enumerate(vocab)}
    om = LM.read_google_ngrams(files, lambda w: ids.get(w, LM.UNKNOWN_WORD_ID))
    for (c, v, t), ov in zip(levels, om.vectors):
        assert [int(x) for x in c] == ov.containers and [int(x) for x in v] == ov.values and t == ov.total

"""Topic-term printing with the reference's output format (pycmf/analysis.py:1-16).

Presentation only, off the fit loop.  Like the reference, each topic lists its 10 heaviest terms in
ascending weight order (the reference hard-codes 10, ignoring `topn_words`; we honour `topn_words`,
whose default is the same 10) and numbers topics from 1.  With `device` given, the per-topic top-k runs
on the GPU (`pycmf_topk_columns`): at the toxic-comments scale the term-topic matrix has 2e5 .. 2e6 rows
and only topn indices per topic come back to the host.
"""
import numpy as np


def top_terms(term_topic_matrix, topn_words, n_topics, device=None):
    """(topics x topn) row indices of the heaviest entries per column, ascending weight (`argsort()[-topn:]`)."""
    M = np.asarray(term_topic_matrix)
    if device is None:
        return np.stack([np.argsort(w, kind="stable")[-topn_words:] for w in M.T[:n_topics]]) if M.shape[1] else \
            np.zeros((0, topn_words), dtype=np.int64)
    from .device import CudaBackend
    dtype = "float32" if M.dtype == np.float32 else "float64"
    be = CudaBackend(device=None if device is True else device, dtype=dtype)
    try:
        F = be.to_device(np.ascontiguousarray(M[:, :n_topics]))
        return be.to_host(be.topk_per_column(F, topn_words)).astype(np.int64)
    finally:
        be.close()


def _topic_lines(term_topic_matrix, idx_to_word, topn_words, n_topics, device=None):
    idx_to_word = np.asarray(idx_to_word)
    picks = top_terms(term_topic_matrix, topn_words, n_topics, device)
    for number, heaviest in enumerate(picks, start=1):
        yield number, ",".join(str(w) for w in idx_to_word[heaviest])


def _print_topic_terms_from_matrix(term_topic_matrix, idx_to_word, topn_words=10, n_topics=100, device=None):
    for number, terms in _topic_lines(term_topic_matrix, idx_to_word, topn_words, n_topics, device):
        print("Topic {}: {}".format(number, terms))


def _print_topic_terms_with_importances_from_matrices(term_topic_matrix, cv_topic_matrix, idx_to_word,
                                                      topn_words=10, n_topics=100, device=None):
    label_weights = np.asarray(cv_topic_matrix).T[:n_topics]
    for (number, terms), weights in zip(_topic_lines(term_topic_matrix, idx_to_word, topn_words, n_topics, device),
                                        label_weights):
        shown = ",".join("{:.3f}".format(x) for x in weights)
        print("Topic {} [{}]: {}".format(number, shown, terms))

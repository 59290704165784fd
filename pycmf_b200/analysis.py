"""Topic-term printing with the reference's output format (pycmf/analysis.py:1-16).

Presentation only, off the fit loop.  Like the reference, each topic lists its 10 heaviest terms in
ascending weight order (the reference hard-codes 10, ignoring `topn_words`; we honour `topn_words`,
whose default is the same 10) and numbers topics from 1.
"""
import numpy as np


def _topic_lines(term_topic_matrix, idx_to_word, topn_words, n_topics):
    idx_to_word = np.asarray(idx_to_word)
    for number, weights in enumerate(np.asarray(term_topic_matrix).T[:n_topics], start=1):
        heaviest = np.argsort(weights)[-topn_words:]
        yield number, ",".join(str(w) for w in idx_to_word[heaviest])


def _print_topic_terms_from_matrix(term_topic_matrix, idx_to_word, topn_words=10, n_topics=100):
    for number, terms in _topic_lines(term_topic_matrix, idx_to_word, topn_words, n_topics):
        print("Topic {}: {}".format(number, terms))


def _print_topic_terms_with_importances_from_matrices(term_topic_matrix, cv_topic_matrix, idx_to_word,
                                                      topn_words=10, n_topics=100):
    label_weights = np.asarray(cv_topic_matrix).T[:n_topics]
    for (number, terms), weights in zip(_topic_lines(term_topic_matrix, idx_to_word, topn_words, n_topics),
                                        label_weights):
        shown = ",".join("{:.3f}".format(x) for x in weights)
        print("Topic {} [{}]: {}".format(number, shown, terms))

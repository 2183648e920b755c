"""CPU oracle for the next-item training/scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package (`seq_recommendations_b200/`) imports this
directory; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may, and there only as the checker / reported CPU baseline, never as the product path.

PARITY UNPINNED.  The reference (efikarra/seq-recommendations) is Python-2 + Keras-2.0.x + Theano; none of
those run in this image, the repository ships no tests, no golden vectors and no saved weights, and it pins no
dependency versions (README.md:5-10).  The arithmetic of the path lives inside Keras/Theano, so this oracle is
a restatement of the Keras-2.0.x layer semantics, anchored on the reference's own
call sites:

  * graph wiring      model.py:241-258 (RNNBaseline), model.py:322-403 (RNNFullModel)
  * batch format      preprocessor.py:16-20, 30-60, 67-94
  * loss / optimizer  experiments_methods.py:41-42  (Adagrad(lr, eps 1e-8, clipnorm 1.), categorical_crossentropy)
  * scoring consumer  model.py:106-112, utils.py:145-178

What IS pinned: the host-side batch format, the synthetic-sequence generator and the likelihood metrics are
checked against the reference's own code (preprocessor.py, sampler.py, utils.py imported with stubbed
third-party modules) through the fixtures under tests/golden/ (generator: tests/golden/make_golden.py).
"""

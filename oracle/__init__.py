"""CPU oracle for the similarity-matrix / hinge-loss / Recall@K hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline -- never as a fallback for the CUDA path.

Pinning status ("parity pinned by execution, not by reference tests"):
the reference ships no tests, golden vectors or known-answer values
(SURVEY.md section 4), so the oracle is pinned the only way available --
``oracle/make_golden.py`` imports the reference's own functions from
``/root/reference`` in the authoring container, runs them on seeded inputs
and commits inputs' seeds + outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks the restatement against those files
on every run (the GPU box has no ``/root/reference``).

Modules
-------
scan_oracle   float64 numpy restatement (cosine, SCAN t2i/i2t, hinge, ranking)
ref_port      float32 torch-CPU port that keeps the reference's per-caption
              loop structure; used as the timed CPU baseline ("port")
ref_loader    imports the unmodified reference with nltk/pycocotools stubs
              (authoring container only)
"""

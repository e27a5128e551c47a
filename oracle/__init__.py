"""CPU oracle for the splashy advect -> project hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker / the timed CPU arm.  The product path (``splashy_b200``) never imports
this package and has no CPU fallback.

Pinning status ("parity unpinned" in the sense of the build contract):
the reference (F#, .NET) cannot be compiled or run in this image and ships no
golden vectors.  The oracle is pinned to the only result pins the reference
holds for this path -- the three still-valid NUnit known-answer tests on
``Pressure.divergence`` (tests/splashy.Tests/Tests.fs:54-73) -- plus the
hand-derived stock-scene values of SURVEY.md App. B, which are re-derived by
``tests/test_oracle_sparse.py`` from the literal transcription in
``sparse_ref.py``.

Modules
-------
sparse_ref   literal, dict-based, fp64 transcription of the reference F#
             (mode "faithful": every quirk Q1-Q11 of SURVEY.md kept).
dense_ref    dense-array restatement with the *intended* (Bridson) semantics
             and the reference's sign / scale / face conventions (Q8, Q9).
dense_ref.c  the same dense restatement in C + OpenMP (timed CPU baseline).
"""

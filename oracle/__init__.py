"""CPU oracle for the falcon clustering hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker or the timed CPU
baseline.  ``falcon_b200`` never imports this package.

What it restates, and how each part is pinned:

=====================  ====================================================  =====================================
stage                  follows                                               pinned by
=====================  ====================================================  =====================================
MurmurHash3_x86_32     sklearn/utils/src/MurmurHash3.cpp:105-157 (public     sklearn KATs (test_murmurhash.py:11-53)
                       algorithm; int key = 4 LE bytes)                      + sklearn.utils.murmurhash3_32 itself
get_dim                /root/reference/falcon/cluster/spectrum.py:172-199    reference function executed by AST
                                                                             extraction -> tests/golden/get_dim.json
binning                /root/reference/falcon/cluster/spectrum.py:250-296    reference ``_to_vector`` executed ->
                       (expression :291, evaluated in float64)               tests/golden/binning.npz
hash + accumulate      SURVEY.md Appendix A.1 (published falcon 0.1.x,       identity with the snapshot's CSR @
                       not in the snapshot)                                  0/1 projection path (spectrum.py:240-246)
buckets / IVF / filter SURVEY.md Appendix A.2 (published falcon 0.1.x +      **parity unpinned** (faiss absent, code
                       faiss IndexIVFFlat semantics; not in the snapshot)    not in snapshot); exhaustive mode is
                                                                             self-checked against brute force
DBSCAN                 sklearn/cluster/_dbscan_inner.pyx:11-41               sklearn ``dbscan_inner`` itself
precursor split        /root/reference/falcon/cluster/cluster.py:334-509     reference functions executed ->
                                                                             tests/golden/postprocess.npz
representatives        SURVEY.md Appendix A.5 (published falcon 0.1.x; the   **parity unpinned** (hand-worked example
                       snapshot keeps a dense descendant, cluster.py:512-553)  in tests/test_oracle_pipeline.py)
=====================  ====================================================  =====================================
"""

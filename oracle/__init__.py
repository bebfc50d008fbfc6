"""CPU oracle for the Stereo R-CNN hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package; the product
(``stereo_rcnn_b200``) never does and fails loudly without its CUDA library.
"""

"""TEST INFRASTRUCTURE ONLY - the parity oracle for the dense-detection hot path.

Nothing under ``oracle/`` is part of the shipped product. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker or as the CPU arm being timed.
The product path (``kuzushiji_vision_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Contents
--------
``torch_oracle``  a torch restatement (CPU and CUDA, fp32) of the reference algorithm:
                  TaskAlignedAssigner, train / inference DFL decode, non_max_suppression.
``ref_loader``    imports the *unmodified* reference from ``/root/reference/yolov12`` when
                  that tree exists (build container only; the GPU box never has it).

Pinning status: the reference's own tests hold no golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself:
``tests/golden/*.npz`` were produced by ``tests/golden/make_golden.py`` from the real
reference functions in the build container, and ``tests/test_oracle_vs_reference.py``
re-runs the live comparison whenever ``/root/reference`` is present.
"""

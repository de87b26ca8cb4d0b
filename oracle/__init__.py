"""TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the rtgs per-ray render path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import or execute it, and there only as the checker / the reported CPU baseline.
The product path (``rt-gaussian-splat-renderer_b200/``) never imports this package.

Modules
-------
ref_numpy   float64 brute-force image oracle (no BVH) + loader activations + orbit camera.
lbvh_ref    integer Morton-30 / sort-key / Karras-hierarchy specification in NumPy.
ref_cpu     ctypes wrapper of ref_cpu.cpp.
ref_cpu.cpp C++/OpenMP restatement with the reference's algorithm shape (K closest-hit
            restarts over a BVH); float = timed CPU baseline, double = large-scene oracle.
taichi_shim a pure-Python stand-in for the subset of Taichi the reference uses, so that the
            reference's OWN source files can be imported from /root/reference in the build
            container to generate the golden vectors committed under tests/golden/.
"""

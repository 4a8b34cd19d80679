"""CPU oracle for the Tuatara OCR hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tuatara_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may call it, and there only as the checker or the timed
CPU baseline -- never as the product path.

What it is: a line-by-line Python restatement of ``/root/reference/tuatara.cpp`` that
calls the *same third-party native kernels* the reference links -- ATen CPU through
``torch`` and OpenCV 4 through ``cv2`` -- at the same call sites (file:line cited on
every function).  The reference itself cannot be built here (``CMakeLists.txt:9`` needs
the OpenCV C++ SDK, which this image lacks; its TorchScript weights need network).

Pinning status: the reference ships NO tests, golden vectors or KATs, so
by the task's definition **parity is unpinned by reference-owned vectors**.  What the
oracle is pinned to instead: (a) cv2 4.13.0 / torch 2.11 CPU executed live, i.e. the very
third-party code the reference delegates all arithmetic to; (b) the numpy / pure-Python
restatements in ``oracle/cvmath.py`` are fuzzed against cv2 in ``tests/test_oracle_*.py``;
(c) the tokenizer table is checked against a gcc-compiled copy of the reference's
constructor logic (``oracle/tokenizer_kat.c`` -- restated, not copied).
"""

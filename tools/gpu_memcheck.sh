#!/bin/bash
# full GPU suite, then compute-sanitizer memcheck over the oracle cases and the full-size custom-filter / caller-buffer / cross-level tests (profiles/r02_memcheck_*)
O=gpurun_out/memcheck; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_against_oracle and (stream or fused)" > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log; tail -4 $O/memcheck.log; grep -c "Invalid\|Illegal" $O/memcheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_reference_fullsize.py -m gpu -x -q -k "custom or layer_a or cross_level" > $O/memcheck2.log 2>&1; echo "memcheck2 rc=$?" >> $O/memcheck2.log; tail -3 $O/memcheck2.log

#!/bin/bash
# 1-GPU call: the continuous off-policy family's parity tests, then the whole GPU suite and smoke()
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_ac_gpu.py -q -m gpu > $O/r02_pytest_ac.txt 2>&1; echo "ac rc=$?"
grep -v "^$" $O/r02_pytest_ac.txt | grep "^FAILED\|^ERROR\|passed\|failed\|Error\|assert\|Mismatch\|Max abs\|Max rel\|^E " | head -60 | cut -c1-260
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_ac_gpu.py > $O/r02_pytest.txt 2>&1; echo "pytest rc=$?"
tail -4 $O/r02_pytest.txt | cut -c1-300
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | cut -c1-300

#!/bin/bash
# Round 2: the whole GPU suite again (the first pass stopped at a too-strict count in tests/test_gpu_bd.py).
tag=${1:-r2t}
mkdir -p gpurun_out
O=gpurun_out/$tag
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_pytest.log 2>&1
echo "pytest exit $?" >> ${O}_pytest.log; tail -6 ${O}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2

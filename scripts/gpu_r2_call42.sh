#!/bin/bash
# round 2, GPU session 42: final tree — whole GPU suite, smoke, default bench line
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
S=r2_s42
timeout 1500 python -m pytest tests -m gpu -q -rA --timeout=600 > $O/${S}_pytest_full.log 2>&1
tail -12 $O/${S}_pytest_full.log > $O/${S}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${S}_smoke.log 2>&1
timeout 1500 python bench.py > $O/${S}_bench_default.log 2> $O/${S}_bench_default.err
tail -3 $O/${S}_pytest.log; tail -3 $O/${S}_smoke.log; head -c 700 $O/${S}_bench_default.log; echo

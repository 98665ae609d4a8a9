#!/bin/bash
# GPU session 34: full GPU suite (incl. full-size faceir / inpaint parity), smoke(), bench lines (own arm + reference arm)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.log | cut -c1-400

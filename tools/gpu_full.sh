#!/bin/bash
# whole GPU suite + smoke + default bench of both arms (what the driver runs at round end)
set -u
O=gpurun_out/${1:-full}
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; tail -4 $O/smoke.log
( time timeout 600 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; tail -c 600 $O/bench_default.err
( time timeout 600 python bench.py --impl reference ) > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 400 $O/bench_ref.err
python tools/bench_summary.py $O/bench_default.json $O/bench_ref.json 2>&1 | tail -20

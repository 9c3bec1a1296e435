#!/bin/bash
# round 2, GPU call A: the whole -m gpu suite (no -x: collect every failure), both bench arms at C2/C3/C4,
# and the ncu launch list of the reference arm.
set -u
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
rm -f gpurun_out/grad_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -rf --durations=15 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
cp gpurun_out/grad_report.jsonl $O/ 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_ours_c2.json 2> $O/bench_ours_c2.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_ref_c2.json 2> $O/bench_ref_c2.err
timeout 300 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_ours_c4.json 2> $O/bench_ours_c4.err
timeout 300 python bench.py --config C4 --impl reference --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_ref_c4.json 2> $O/bench_ref_c4.err
timeout 300 python bench.py --config C3 --train --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_ours_c3.json 2> $O/bench_ours_c3.err
timeout 300 python bench.py --config C3 --train --impl reference --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_ref_c3.json 2> $O/bench_ref_c3.err
timeout 300 python bench.py --views 1 --streams 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_ours_v1.json 2> $O/bench_ours_v1.err
# ncu launch list of the reference arm (its renderCUDA, CUB sort, torch bmm's): plain run first
timeout 300 python bench.py --impl reference --steps 1 --warmup 3 --views 1 --no-cpu-baseline > $O/ref_plain.json 2> $O/ref_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 500 --csv --log-file $O/ref_ncu_launches.csv \
    python bench.py --impl reference --steps 1 --warmup 3 --views 1 --no-cpu-baseline > $O/ref_ncu.log 2>&1
tail -5 $O/pytest.log

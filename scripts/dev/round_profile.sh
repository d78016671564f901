#!/bin/bash
# One GPU-box pass that produces every artifact of a round: tests, smoke, both bench arms, the 03w bench, the ncu launch lists
# of ONE training step (cfg01 and 03w) and the ncu --set full captures of the layer / head / wide kernels (each only after the
# same command has exited 0 without ncu).  Every command under its own timeout.
#   gpurun --timeout 1700 -- 'bash scripts/dev/round_profile.sh r02f'
tag=${1:-rXX}
out=gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
timeout 400 python bench.py --impl reference > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
timeout 500 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 300 python bench.py --workload 03w --steps 5 --warmup 3 --sustained-steps 100 --no-cpu-baseline > $out/${tag}_bench_03w.json 2> $out/${tag}_bench_03w.err
# launch lists of exactly one training step
timeout 300 python scripts/dev/step_profile_any.py 01 > /dev/null 2>&1 || exit 1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_step_01.csv python scripts/dev/step_profile_any.py 01 > $out/${tag}_ncu1.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_step_03w.csv python scripts/dev/step_profile_any.py 03w > $out/${tag}_ncu1w.log 2>&1
# full captures: the cfg01 layer / head kernels, the wide kernels
cmd="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-decode"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'layer_bwd_db_kernel|layer_(bwd|fwd)_tc_kernel|head_(bwd|fwd)_tc_kernel|video_conv_tc_kernel' -s 26 -c 16 -o $out/${tag}_prof $cmd > $out/${tag}_ncu2.log 2>&1
ncu -i $out/${tag}_prof.ncu-rep --page raw --csv > $out/${tag}_layer_kernels_ncu_full_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'wide_gemm_kernel|wide_wgrad_kernel' -s 10 -c 12 -o $out/${tag}_prof_wide python scripts/dev/wide_layer_times.py > $out/${tag}_ncu2w.log 2>&1
ncu -i $out/${tag}_prof_wide.ncu-rep --page raw --csv > $out/${tag}_wide_kernels_ncu_full_raw.csv 2>/dev/null
ls -la $out | tail -14

#!/bin/bash
# One GPU-box pass that produces every artifact of a round: tests, smoke, both bench arms, the ncu launch list and the
# ncu --set full capture of the layer / head kernels (each only after the same command has exited 0 without ncu).
#   gpurun --timeout 1500 -- 'bash scripts/dev/round_profile.sh r01c'
tag=${1:-rXX}
out=gpurun_out
set -x
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
python bench.py --impl reference > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
cmd="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-decode"
$cmd > $out/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $cmd > $out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'layer_(bwd|fwd)_tc_kernel|head_(bwd|fwd)_tc_kernel' -s 24 -c 14 -o $out/${tag}_prof $cmd > $out/${tag}_ncu2.log 2>&1
ncu -i $out/${tag}_prof.ncu-rep --page raw --csv > $out/${tag}_raw.csv 2>/dev/null
ls -la $out | tail -12

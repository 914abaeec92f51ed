#!/bin/bash
# (one step = 24 kernel launches incl. the 5 of the CUB sort: -c 72 = three steps, -c 24 = one)
# The round's evidence in one GPU call: GPU tests, the default bench line (+ reference arm), the ncu launch list and a metric-list
# capture of one step of the same command.  usage: tools/final_profile.sh <tag>   (outputs under gpurun_out/)
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python tools/show_bench.py gpurun_out/${tag}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_reference.json 2> gpurun_out/${tag}_reference.err; tail -c 600 gpurun_out/${tag}_reference.json
CMD="python bench.py --steps 3 --warmup 5 --no-cpu-baseline --no-extras --min-seconds 0"
$CMD > gpurun_out/${tag}_plain.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 72 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
M=$(python -c "import sys; sys.path.insert(0,'tools'); import make_profiles as m; print(','.join(m.METRICS))")
ncu --metrics $M --clock-control none -s 100 -c 24 -o gpurun_out/${tag}_metrics -f $CMD > gpurun_out/${tag}_ncu2.log 2>&1
ls -la gpurun_out/${tag}_*

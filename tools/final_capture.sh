# One gpurun call that re-captures every committed artefact of the headline on the current sources:
#   GPU test suite, A/B of string pages against a previous build (copy it to strawboat_b200/csrc/libsb_old.so first;
#   the step just fails without it), ncu --set full capture -> profiles/r2_ncu_traffic.json + summary csv, launch list,
#   and the default `python bench.py` line.  Everything lands in gpurun_out/final7/; copy what should be judged to profiles/.
# usage: gpurun --timeout 780 -- 'bash tools/final_capture.sh'
set -x
O=gpurun_out/final7; mkdir -p $O
(timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > $O/pytest.log 2>&1; cat $O/pytest.log
timeout 200 python tools/ab_strings.py libsb_old.so libstrawboat_b200.so > $O/ab.log 2>&1; tail -3 $O/ab.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sb_ --launch-skip 9 -c 6 -f -o $O/r2_full python bench.py --steps 2 --warmup 3 --no-extras --no-cpu --no-e2e > $O/ncu_full.log 2>&1
python tools/ncu_traffic.py $O/r2_full.ncu-rep > $O/traffic.log 2>&1; cp profiles/r2_ncu_traffic.json $O/
python tools/ncu_summary.py $O/r2_full.ncu-rep > $O/r2_ncu_full_config2.csv 2>> $O/traffic.log
rm -f $O/r2_full.ncu-rep
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_config2.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > /dev/null 2>&1
timeout 500 python bench.py > $O/r2_bench_config2.json 2> $O/bench.err; tail -c 600 $O/r2_bench_config2.json

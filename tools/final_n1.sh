set -x
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_final_pytest.log 2>&1; tail -3 gpurun_out/r2_final_pytest.log
timeout 900 python bench.py > gpurun_out/r2_final2_n1.json 2> gpurun_out/r2_final2_n1.err; tail -c 600 gpurun_out/r2_final2_n1.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final2_ref.json 2> gpurun_out/r2_final2_ref.err; tail -c 400 gpurun_out/r2_final2_ref.json
timeout 300 python bench.py --steps 3 --warmup 3 --profile > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final2_launches.csv python bench.py --steps 3 --warmup 3 --profile > gpurun_out/ncu_final2.log 2>&1
tail -2 gpurun_out/ncu_final2.log

set -x
cd $GRAFT_REPO_ROOT
T=/tmp/ncu_r02; mkdir -p $T gpurun_out/ev
ncu --set full --clock-control none --import-source on -k regex:"scan|slice" -o $T/r02_k1 python scripts/profile_all_kernels.py > gpurun_out/ev/r02_k1.log 2>&1
python scripts/ncu_summarize.py $T/r02_k1.ncu-rep gpurun_out/ev/r02_k1
ncu --set full --clock-control none --import-source on -k regex:"classify|live_|modeac|dc_|float_block|convert_kernel|crc_batch" -o $T/r02_rest python scripts/profile_all_kernels.py > gpurun_out/ev/r02_rest.log 2>&1
python scripts/ncu_summarize.py $T/r02_rest.ncu-rep gpurun_out/ev/r02_rest
ncu --set full --clock-control none --import-source on -k regex:"scan2|slice|classify_warp" -c 12 -o gpurun_out/ev/r02_pipeline python scripts/profile_pipeline.py 60 uc8 > gpurun_out/ev/r02_pipeline.log 2>&1
python scripts/ncu_summarize.py gpurun_out/ev/r02_pipeline.ncu-rep gpurun_out/ev/r02_pipeline
ncu --cache-control none --clock-control none --metrics lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:"scan2|slice|classify_warp|live_" -c 20 --csv --log-file gpurun_out/ev/r02_pipeline_cache_control_none.csv python scripts/profile_pipeline.py 60 uc8 > gpurun_out/ev/r02_pipeline_ccn.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev/r02_ncu_launches_bench_steps2_warmup1.csv python bench.py --steps 2 --warmup 1 --other-configs none > gpurun_out/ev/r02_bench_under_ncu.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/ev/r02_final_bench.json 2> gpurun_out/ev/r02_final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ev/r02_final_bench_reference.json 2> gpurun_out/ev/r02_final_bench_reference.err
python -m pytest tests -m gpu -q > gpurun_out/ev/r02_final_pytest.log 2>&1
tail -2 gpurun_out/ev/r02_final_pytest.log
ls -la gpurun_out/ev

set -x
python bench.py > gpurun_out/r3_bench_n1.json 2> gpurun_out/r3_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r3_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/r3_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_fast_kernel -s 2 -c 1 -o gpurun_out/r3_asm5792 python bench.py --steps 2 --warmup 3 --skip-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/r3_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_tma_kernel -s 5 -c 1 -o gpurun_out/r3_spmv5792 python bench.py --steps 2 --warmup 3 --skip-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/r3_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"assemble_fast_kernel|cell_setup_damage" -s 18 -c 3 -o gpurun_out/r3_dmg1448 python tools/r2_probe.py dmg100 > gpurun_out/r3_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_fast_kernel -s 14 -c 1 -o gpurun_out/r3_p1_2896 python tools/r2_probe.py p1 > gpurun_out/r3_ncu5.log 2>&1

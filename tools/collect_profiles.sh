#!/bin/bash
# Round-2 evidence run on ONE B200 (gpurun): bench lines, reference arm, launch list, ncu summaries, GEMM table.
# Everything lands in gpurun_out/r02/ ; copy what should be judged into profiles/.
set -u
O=${1:-gpurun_out/r02}
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py --steps 50 --warmup 5 --batch 16 --no-library-bar > $O/bench_b16.json 2> $O/bench_b16.err
python bench.py --steps 20 --warmup 5 --neurons 512 --no-library-bar --no-cpu-baseline > $O/bench_n512.json 2> $O/bench_n512.err
python bench.py --steps 20 --warmup 5 --eval-mode --no-library-bar --no-cpu-baseline > $O/bench_evalmode.json 2> $O/bench_evalmode.err
python bench.py --steps 6 --warmup 3 --workload scaled --no-library-bar --no-cpu-baseline > $O/bench_scaled_config5.json 2> $O/bench_scaled.err
python bench.py --steps 32 --warmup 5 --workload multisession --no-library-bar --no-cpu-baseline > $O/bench_multisession_config4.json 2> $O/bench_multi.err
python tools/gemm_bench.py > $O/gemm_bench.txt 2>&1
python tools/attn_bench.py 256 1 > $O/attn_bench.txt 2>&1
python tools/attn_bench.py 256 0 >> $O/attn_bench.txt 2>&1
python tools/attn_bench.py 16 1 16 64 1000 >> $O/attn_bench.txt 2>&1
echo '# causal mask (mask-aware block skipping; FLOPs counted for the full S x S product)' >> $O/attn_bench.txt
python tools/attn_bench.py 16 1 16 64 1000 2 >> $O/attn_bench.txt 2>&1
./tools/micro/umma_rate > $O/umma_rate.txt 2>&1
python tools/step_profile.py > $O/step_profile.txt 2>&1
# launch list of one step (serialised, cold cache: shares only)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-bar --sustained-seconds 0 > $O/plain_launch.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1500 -c 330 --csv \
    --log-file $O/ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-bar --sustained-seconds 0 > $O/ncu_launch.log 2>&1
# full captures of the hot kernels
python tools/gemm_one.py qkv > $O/one_qkv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_ts -s 3 -c 1 -o $O/prof_gemm_ts_qkv python tools/gemm_one.py qkv > $O/ncu_qkv.log 2>&1
python tools/gemm_one.py mulaux > $O/one_mulaux.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_ts -s 3 -c 1 -o $O/prof_gemm_ts_mulaux python tools/gemm_one.py mulaux > $O/ncu_mulaux.log 2>&1
python tools/attn_bench.py 256 1 > $O/one_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_pipe -s 3 -c 1 -o $O/prof_attn_fwd python tools/attn_bench.py 256 1 > $O/ncu_attn_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_ws -s 3 -c 1 -o $O/prof_attn_bwd python tools/attn_bench.py 256 1 > $O/ncu_attn_bwd.log 2>&1
ls -la $O | head -50

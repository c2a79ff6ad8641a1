#!/bin/bash
# round-2 profiling evidence: launch list of the default bench command, ncu --set full of the dominant kernels
mkdir -p gpurun_out/ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
K='regex:blstm|gemm_tc|stft_kernel|feature_|mask_istft|activity_kernel|median_|segments_|fold_|cast_bf16|pack_whh|head_expand|instance_norm'
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config3 --no-parity > gpurun_out/r2_launches_bench.out 2> gpurun_out/r2_launches_bench.err
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2_launches.csv)"
full() {  # name, kernel regex, skip, command...
  local name=$1 kre=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s "$skip" -c 1 -f -o gpurun_out/ncu/$name "$@" > gpurun_out/ncu/$name.log 2>&1
  python scripts/ncu_summary.py gpurun_out/ncu/$name.ncu-rep gpurun_out/r2_ncu_$name.txt > /dev/null 2>> gpurun_out/ncu/$name.log
  echo "$name rc=$? $(grep -E 'gpu__time_duration|dram__bytes' gpurun_out/r2_ncu_$name.txt | tr '\n' ' ')"
}
full rec_ts_416rows   blstm_rec_ts 2 python scripts/profile_rec.py --rows 416 --clusters 32 --tiles 2 --subs 2 --frames 2000 --reps 1
full rec_ts_104rows   blstm_rec_ts 2 python scripts/profile_rec.py --rows 104 --clusters 8 --tiles 2 --subs 1 --frames 4000 --reps 1
full rec_ts_208rows_pingpong blstm_rec_ts 2 python scripts/profile_rec.py --rows 208 --clusters 16 --tiles 2 --subs 2 --frames 4000 --reps 1
full rec_ts_8rows_1tile blstm_rec_ts 2 python scripts/profile_rec.py --rows 8 --clusters 8 --tiles 1 --subs 1 --frames 4000 --reps 1
full rec_ts_832rows_pingpong blstm_rec_ts 2 python scripts/profile_rec.py --rows 832 --clusters 64 --tiles 2 --subs 2 --frames 2000 --reps 1
full blstm_bwd        blstm_bwd 4 python bench.py --config c5 --steps 1 --warmup 1 --no-cpu-baseline
full stft             stft_kernel 2 python scripts/profile_frontend.py --meetings 8 --reps 1
full feature_stats    feature_stats 2 python scripts/profile_frontend.py --meetings 8 --reps 1
full feature_write    feature_write 2 python scripts/profile_frontend.py --meetings 8 --reps 1
full mask_istft       mask_istft_1024 2 python scripts/profile_istft.py --meetings 4 --reps 1
full gemm_head        gemm_tc 2 python scripts/profile_gemm.py --meetings 2 --only head --reps 1
full gemm_b1_in       gemm_tc 2 python scripts/profile_gemm.py --meetings 2 --only b1_in --reps 1
python scripts/profile_frontend.py --meetings 16 > gpurun_out/r2_frontend_microbench.txt 2>&1; cat gpurun_out/r2_frontend_microbench.txt
python scripts/profile_gemm.py --meetings 8 > gpurun_out/r2_gemm_microbench.txt 2>&1; cat gpurun_out/r2_gemm_microbench.txt
ls -la gpurun_out/ncu/*.ncu-rep | awk '{print $5, $9}'
# keep the two recurrence captures, drop the other raw reports (64 MiB merge limit)
find gpurun_out/ncu -name '*.ncu-rep' ! -name 'rec_ts_416rows*' ! -name 'rec_ts_8rows*' ! -name 'gemm_head*' -delete
du -sh gpurun_out

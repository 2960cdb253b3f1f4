#!/bin/bash
# Round-2 profile captures (run under gpurun on one B200; every ncu run follows a plain run of the same command that exited 0).
# Outputs under gpurun_out/r02prof/; tools/profile_r02_summary.py turns them into the text / JSON files committed under profiles/.
set -u
O=gpurun_out/r02prof
mkdir -p $O
BENCH="python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-sampling --no-rvq --no-full-step"
# 1. launch list of the end of the eager run of the bench workload (cold-cache, serialised: compare SHARES).  Five steps launch ~7200
#    kernels (1670 in the first, ~1382 in each of the others); the first 5500 are skipped and the summary keeps the last step: from the
#    last add_noise kernel, which runs once per step, to the end.  (Either way ncu intercepts every launch: this pass takes ~17
#    GPU-minutes on its own -- tools/profile_r02_launchlist.sh runs just it, tools/profile_r02_partial.sh just a few full captures.)
$BENCH > $O/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 5500 -c 30000 --csv --log-file $O/launches_all.csv $BENCH > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
# 2. full captures of the dominant kernels
cap() {  # name, regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/${name}_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $O/$name "$@" > $O/${name}_ncu.log 2>&1
  echo "$name rc=$?"
}
cap gemm_pair_6016x1280x10240 gemm_kernel 2 1 python tools/gemm_one.py 6016 1280 10240 1 1 0 256 3
cap gemm_auto_6016x1280x10240 gemm_kernel 2 1 python tools/gemm_one.py 6016 1280 10240 1 1 0 0 3
cap gemm_shortk_24064x2560x320 gemm_kernel 2 1 python tools/gemm_one.py 24064 2560 320 1 1 0 0 3
cap gemm_wgrad_1280x1280x6016 gemm_kernel 2 1 python tools/gemm_one.py 1280 1280 6016 0 0 2 0 3
cap attn_d40_self attn_kernel 3 3 python tools/attn_one.py full_d40_self 2
cap rvq_tc rvq_encode_tc 1 1 python tools/rvq_tc_one.py 256
cap rvq_rest "rvq_(encode_kernel|decode)" 3 3 python tools/rvq_one.py 32
cuobjdump -sass prompt_tts_b200/libpt_b200.so | grep -oE "UTCHMMA[.A-Z0-9_]*|LDTM[.A-Za-z0-9_]*|STTM[.A-Za-z0-9_]*|UTMALDG[.A-Z0-9_]*|UTMASTG[.A-Z0-9_]*|UTCBAR[.A-Z0-9_]*|REDG[.A-Za-z0-9_]*|HMMA[.A-Z0-9_]*" | sort | uniq -c > $O/sass_mnemonics.txt
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
ls -la $O | head -40

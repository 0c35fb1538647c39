#!/usr/bin/env bash
# tools/r2_session.sh <steps...> — round-2 GPU sessions: each named step under its own timeout, its output in gpurun_out/<name>.log
# as soon as it is done. Nothing here is a bench value taken under a profiler.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p "$OUT"
step() { local secs=$1 name=$2; shift 2; local t0=$SECONDS
    timeout --signal=TERM --kill-after=20 "$secs" "$@" > "$OUT/$name.log" 2>&1
    echo "$name rc=$? $((SECONDS - t0))s" | tee -a "$OUT/session_steps.txt"; return 0; }
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > "$OUT/gpu.txt" 2>&1
free -g > "$OUT/host.txt"; nproc >> "$OUT/host.txt"
for s in "$@"; do
  case "$s" in
    fullcfg)   step 900 fullcfg python -m pytest tests/test_full_config_gpu.py -q -x -s -rs ;;
    probe)     step 120 fusion_probe tools/microbench/_build/fusion_probe 1000 ;;
    trace)     step 120 mega_trace python tools/mega_trace.py ;;
    trace_ll)  step 120 mega_trace_ll python tools/mega_trace.py --ll ;;
    trace_v2)  step 120 mega_trace_v2 python tools/mega_trace.py --v2 ;;
    trace_v2f) step 120 mega_trace_v2f python tools/mega_trace.py --v2 --fuse-down ;;
    trace_v2fc) step 120 mega_trace_v2fc python tools/mega_trace.py --v2 --fuse-down --calibrate ;;
    int8tests) step 600 int8tests python -m pytest tests/test_engine_gpu.py -q -x -k "int8 or INT8 or wd2 or 2-ll or golden" ;;
    benchtp*)  n=${s#benchtp}; step 600 "bench_tp${n}" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus "$n" --steps 20 --warmup 5; grep -h '^{' "$OUT/bench_tp${n}.log" | tail -1 > "$OUT/bench_tp${n}.json" ;;
    trace_tiny) step 200 mega_trace_tiny_v2 python tools/mega_trace.py --config tinyllama-1.1b --layers 22 --pos 1700 --v2 --calibrate
               step 200 mega_trace_tiny_v1 python tools/mega_trace.py --config tinyllama-1.1b --layers 22 --pos 1700
               step 200 mega_trace_110m_v2 python tools/mega_trace.py --config stories110M --wdtype f32 --kvdtype f32 --layers 12 --pos 128 --v2 ;;
    widetests) step 600 widetests python -m pytest tests/test_engine_gpu.py -q -x -k "wide_heads or int8 or medium" ;;
    tcbatch)   step 600 tcbatch_tests python -m pytest tests/test_zz_batch_gpu.py -q -x -s -k "tensor_core"
               step 600 batch_bench_tc python tools/batch_bench.py --variants tc,tc+graph --tc-batches 8,16,32,64 --json ;;
    pftests)   step 900 pftests python -m pytest tests/test_prefill_gpu.py -q -x ;;
    batchtests) step 900 batchtests python -m pytest tests/test_zz_batch_gpu.py -q -x ;;
    tcbench)   step 600 batch_bench_tc python tools/batch_bench.py --variants tc+graph --tc-batches 8,16,32,64 --json ;;
    caltest)   step 300 caltest python -m pytest tests/test_engine_gpu.py -q -k "calibrated" ;;
    debug_v2)  step 240 mega_debug_v2 python tools/mega_debug.py --v2 ;;
    sweep)     step 400 mega_sweep python tools/mega_sweep.py ;;
    libab)     for v in "" _phsmem _nosp ""; do step 300 "libab${v:-_base}_$RANDOM" env SLLM_LIB=$PWD/simplellminference_b200/lib/libsllm_b200$v.so python tools/mega_sweep.py --variants v2f+cal --debug 0,0; done ;;
    sweep_small) step 300 sweep_tiny python tools/mega_sweep.py --config tinyllama-1.1b --pos 1700 --variants v1,v1+cal,v1f,v2,v2+cal,v2f,v2f+cal --debug 0
               step 300 sweep_110m python tools/mega_sweep.py --config stories110M --wdtype f32 --kvdtype f32 --pos 128 --variants v1,v1+cal,v2,v2+cal --debug 0
               step 300 sweep_8b python tools/mega_sweep.py --config llama3-8b --pos 7800 --variants v1,v2f,v2f+cal --debug 0 ;;
    enginetests) step 900 enginetests python -m pytest tests/test_engine_gpu.py -q -s ;;
    steal)     step 300 case_v2fs_7b2 python tools/sanitize_case.py v2fs_7b2
               step 300 case_v2fs_7b2s python tools/sanitize_case.py v2fs_7b2s
               step 300 sweep_steal python tools/mega_sweep.py --variants v2f+cal,v2fs+cal,v2f+cal,v2fs+cal --debug 0,0
               step 200 mega_trace_steal python tools/mega_trace.py --v2 --fuse-down --calibrate --steal ;;
    sweep_ab)  step 400 mega_sweep_ab python tools/mega_sweep.py --variants v2f+cal --debug 0,4,8,12,0 ;;
    san_*)     c=${s#san_}; tool=${c%%:*}; case_=${c#*:}; step 600 "sanitize_${tool}_${case_}" env SLLM_COMPARE=0 compute-sanitizer --tool "$tool" --print-limit 30 python tools/sanitize_case.py "$case_" ;;
    case_*)    step 300 "case_${s#case_}" python tools/sanitize_case.py "${s#case_}" ;;
    ncu_v2f)   PS="python tools/profile_step.py --v2 --fuse-down --calibrate --pos 520 --steps 2"
               step 120 ncu_v2f_plain $PS
               step 600 ncu_v2f_full ncu --set full --clock-control none --import-source on -k regex:mega2_step -s 30 -c 1 -o "$OUT/r02_mega2_v2f" -f $PS ;;
    ncu_bench) step 600 ncu_bench_launches ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'mega|pf_|embed|rmsnorm' -c 3000 --csv --log-file "$OUT/r02_bench_launches.csv" python bench.py --steps 20 --warmup 5 --no-cpu-baseline ;;
    tp*)       n=${s#tp}; step 900 "tp${n}_check" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29600 + n)) tests/tp_check.py ;;
    cfgs)      step 300 bench_cfg2 python bench.py --config stories110M --wdtype f32 --kvdtype f32 --prompt-len 128 --no-cpu-baseline
               step 300 bench_cfg3_bf16 python bench.py --config tinyllama-1.1b --prompt-len 1700 --no-cpu-baseline
               step 300 bench_cfg3_int8 python bench.py --config tinyllama-1.1b --wdtype int8 --prompt-len 1700 --no-cpu-baseline
               step 400 bench_cfg5 python bench.py --config llama3-8b --prompt-len 7800 --steps 64 --no-cpu-baseline
               step 300 bench_7b_int8 python bench.py --wdtype int8 --no-cpu-baseline
               for n in cfg2 cfg3_bf16 cfg3_int8 cfg5 7b_int8; do grep -h '^{' "$OUT/bench_$n.log" | tail -1 > "$OUT/bench_$n.json"; done ;;
    batchb)    step 400 batch_bench python tools/batch_bench.py --batches 1,4,8,16 --exp-batches 8,16 --variants plain,graph,graph+rows4,graph+rows4+ksplit --json ;;
    ncu_pf)    PB="python tools/prefill_bench.py --tokens 512 --reps 2"
               step 120 ncu_pf_plain $PB
               step 600 ncu_pf_gemm ncu --set full --clock-control none --import-source on -k regex:pf_gemm_kernel -s 300 -c 8 -o "$OUT/r02_pf_gemm" -f $PB
               step 400 ncu_pf_attn ncu --set full --clock-control none --import-source on -k regex:pf_attn_kernel -s 40 -c 2 -o "$OUT/r02_pf_attn" -f $PB ;;
    ncu_tc)    TB="python tools/batch_bench.py --variants tc --tc-batches 16 --context 520 --steps 2 --json"
               step 300 ncu_tc_plain $TB
               step 600 ncu_tc_gemm ncu --set full --clock-control none --import-source on -k regex:pf_gemm_kernel -s 66800 -c 8 -o "$OUT/r02_tc_gemm" -f $TB
               step 600 ncu_tc_mha ncu --set full --clock-control none --import-source on -k regex:mha_paged -s 16660 -c 2 -o "$OUT/r02_tc_mha" -f $TB ;;
    trace_tp*) n=${s#trace_tp}; step 300 "mega_trace_ll_tp${n}" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29800 + n)) tools/mega_trace.py --ll --layers 8 ;;
    smoke)     step 600 smoke python -c 'import __graft_entry__ as g; g.smoke(); print("__SMOKE_OK__")' ;;
    bench_ref) step 600 bench_ref python bench.py --impl reference --steps 20 --warmup 5; grep -h '^{' "$OUT/bench_ref.log" | tail -1 > "$OUT/bench_ref.json" ;;
    bench_drv) step 600 bench_drv python bench.py --gpus 1 --steps 20 --warmup 5; grep -h '^{' "$OUT/bench_drv.log" | tail -1 > "$OUT/bench_drv.json" ;;
    pprobe)    step 120 prefetch_probe tools/microbench/_build/prefetch_probe ;;
    cprobe)    step 120 consumer_probe tools/microbench/_build/consumer_probe ;;
    v2tests)   step 900 v2tests python -m pytest tests/test_engine_gpu.py -q -k "v2" ;;
    v2check)   step 300 v2check python tools/v2_check.py ;;
    bench_v2)  step 600 bench_v2 python bench.py --no-batch --no-cpu-baseline --mega-v2; grep -h '^{' "$OUT/bench_v2.log" | tail -1 > "$OUT/bench_v2.json" ;;
    bench_v2f) step 600 bench_v2f python bench.py --no-batch --no-cpu-baseline --mega-v2 --mega-fuse-down; grep -h '^{' "$OUT/bench_v2f.log" | tail -1 > "$OUT/bench_v2f.json" ;;
    debug)     step 240 mega_debug python tools/mega_debug.py ;;
    pytest)    step 1200 pytest_gpu python -m pytest tests -q -m gpu -rxXs --deselect tests/test_full_config_gpu.py ;;
    pytest_all) step 1500 pytest_gpu python -m pytest tests -q -m gpu -rxXs ;;
    bench)     step 600 bench python bench.py --no-batch; grep -h '^{' "$OUT/bench.log" | tail -1 > "$OUT/bench.json" ;;
    bench_full) step 900 bench python bench.py; grep -h '^{' "$OUT/bench.log" | tail -1 > "$OUT/bench.json" ;;
    *) step 900 "custom_$(echo "$s" | tr -c 'a-zA-Z0-9' '_' | cut -c1-40)" bash -c "$s" ;;
  esac
done
echo "session done" | tee -a "$OUT/session_steps.txt"

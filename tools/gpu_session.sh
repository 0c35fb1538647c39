#!/usr/bin/env bash
# tools/gpu_session.sh — everything the FIRST GPU call of a session should bring back, in order of importance, each step
# under its own timeout and with its output in gpurun_out/ as soon as it is done (a call that is cut off keeps what
# finished). Nothing here is a bench value taken under a profiler: ncu runs only after the same command exited 0 without it.
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh'            (one GPU)
#   bash tools/gpu_session.sh quick                                  (tests of the batched decode + bench only)
#   gpurun --gpus 8 --timeout 3000 -- 'bash tools/gpu_session.sh'   (adds tensor-parallel parity + bench at 2 / 4 / 8 GPUs at the end)
#
# Order: 1 the never-executed batched-decode tests (progressive log: tests/batch_check.py), 2 the bench line,
# 3 batched throughput with every experimental variant, 4 the whole GPU suite, 5 per-phase timeline of the decode
# megakernel + the fusion probe (reductions vs barrier, clusters, L2 vs HBM ingest: DESIGN.md section 10), 6 launch list + ncu --set full of the batched GEMV and the paged attention kernel.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p "$OUT"
MODE="${1:-full}"
step() {   # step <seconds> <name> <command...>: run under a timeout, log, never stop the script
    local secs=$1 name=$2; shift 2
    local t0=$SECONDS
    timeout --signal=TERM --kill-after=20 "$secs" "$@" > "$OUT/$name.log" 2>&1
    local rc=$?
    echo "$name rc=$rc $((SECONDS - t0))s" | tee -a "$OUT/session_steps.txt"
    return 0
}
: > "$OUT/session_steps.txt"
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > "$OUT/gpu.txt" 2>&1

step 600 batch_check      python tests/batch_check.py "$OUT/batch_check_progress.log"
step 900 bench            python bench.py
grep -h '^{' "$OUT/bench.log" | tail -1 > "$OUT/bench.json"
step 300 batch_bench      python tools/batch_bench.py --batches 1,2,4,8,16,32 --exp-batches 4,8,16,32 \
                          --variants plain,graph,graph+rows4,graph+rows4+ksplit --json
[ "$MODE" = quick ] && exit 0
step 1500 pytest_gpu      python -m pytest tests -q -m gpu -rxXs --junitxml="$OUT/pytest_gpu.xml"
step 180 mega_trace       python tools/mega_trace.py
step 180 fusion_probe     tools/microbench/_build/fusion_probe 2000
# experimental megakernel with the down projection fused into the gate_up phase: parity first, then the same bench line and timeline
step 600 fuse_check       python tests/fuse_check.py "$OUT/fuse_check_progress.log"
step 420 bench_fuse       python bench.py --mega-fuse-down --no-batch --no-cpu-baseline
step 180 mega_trace_fuse  python tools/mega_trace.py --fuse-down
# ---- profiler passes (after the plain commands above have run)
PS="python tools/profile_step.py --layers 2 --pos 512 --steps 2"
step 120 batch8_plain     $PS --batch 8
step 300 batch8_launches  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/batch8_launches.csv" $PS --batch 8
step 600 batch8_ncu_gemv  ncu --set full --clock-control none --import-source on -k regex:bgemv -s 40 -c 6 -o "$OUT/batch8_bgemv" -f $PS --batch 8
step 400 batch8_ncu_mha   ncu --set full --clock-control none --import-source on -k regex:mha_paged -s 8 -c 2 -o "$OUT/batch8_mha_paged" -f $PS --batch 8
# ---- more than one GPU on the box (gpurun --gpus N): tensor-parallel parity and the bench line at every power of two up to N
NGPU=$(nvidia-smi -L 2>/dev/null | wc -l)
for n in 2 4 8; do
    [ "$NGPU" -ge "$n" ] || break
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n))"
    step 900 "tp${n}_check" $TR tests/tp_check.py
    step 600 "tp${n}_bench" $TR bench.py --gpus "$n"
    grep -h '^{' "$OUT/tp${n}_bench.log" | tail -1 > "$OUT/tp${n}_bench.json"
done
echo "session done" | tee -a "$OUT/session_steps.txt"

# usage: bash tools/gpu_trace.sh "<op indices>"  -> gpurun_out/trace.log (role timelines of CTA 0)
export B2D_TRACE=1 B2D_TRACE_DUMP=1 B2D_LIB=tools/ubench/build/libb2det_trace.so
for op in $1; do
timeout 100 python tools/one_op.py --op $op --reps 1 2>&1 | tail -30
done > gpurun_out/trace.log 2>&1

python tools/one_op.py --op 3 4 10 11 --reps 1 > gpurun_out/oneop_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_tc -s 89 -c 4 --csv --log-file gpurun_out/ncu_ops_3_4.csv python tools/one_op.py --op 3 4 10 11 --reps 1 > gpurun_out/oneop_ncu.log 2>&1
echo "ncu rc=$?"
B2D_L2_FETCH=32 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_tc -s 89 -c 4 --csv --log-file gpurun_out/ncu_ops_3_4_l2f32.csv python tools/one_op.py --op 3 4 10 11 --reps 1 > gpurun_out/oneop_ncu2.log 2>&1
echo "ncu2 rc=$?"
ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_tc -s 89 -c 4 --csv --log-file gpurun_out/ncu_ops_3_4_warm.csv python tools/one_op.py --op 3 4 10 11 --reps 1 > gpurun_out/oneop_ncu3.log 2>&1
echo "ncu3 rc=$?"

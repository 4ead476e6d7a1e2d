#!/bin/bash
# round-2 ncu capture of the HBM-side (sampler) kernels of the native loops + CUDA-event stage timings
mkdir -p gpurun_out
python tools/profile_hbm_kernels.py > gpurun_out/r2_hbm_kernels_events.json 2> gpurun_out/r2_hbm_events.err; echo "events rc=$?"; cat gpurun_out/r2_hbm_kernels_events.json
CSL_PROFILE_SHORT=1 ncu --set full --clock-control none --import-source on \
  -k regex:'k_stretch_propose_prepare|k_stretch_combine_accept|k_mh_propose_prepare|k_mh_combine_accept|k_prepare|k_emcee_rows' \
  --launch-skip 4 -c 16 -f -o /tmp/r2_hbm_kernels python tools/profile_hbm_kernels.py > gpurun_out/ncu_hbm.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r2_hbm_kernels.ncu-rep --page raw --csv > /tmp/r2_hbm_raw.csv
python tools/summarize_ncu_hbm.py /tmp/r2_hbm_raw.csv gpurun_out/r2_ncu_full_hbm_kernels; cat gpurun_out/r2_ncu_full_hbm_kernels.txt

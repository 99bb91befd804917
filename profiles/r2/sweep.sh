# How the sweep_*.txt / split_*.txt files of this directory were produced (one gpurun call each):
# environment switches of the engine (csrc/engine.cu: cwtb_create) or compile-time variants
# (profiles/micro/build_variant.py TAG -D...) around the two timing scripts
#   profiles/micro/split_timing.py [lib.so ...]     config-2 step and its parts (SPLIT_PARTS=full,exact,...)
#   profiles/micro/config_kernels.py 2,3,4,5 [--prof]  device-timed step of every configuration
# e.g.
#   for p in 0 1 2; do CWTB_PRIO=$p python profiles/micro/split_timing.py; done          # split_c_priority.txt
#   for m in 0 1 2; do CWTB_DENSE_MARGIN=$m SPLIT_PARTS=full,exact python profiles/micro/split_timing.py; done
#   for ch in 2 4; do for g in 2 3 4 8 16; do CWTB_CHAINS=$ch CWTB_GROUP=$g ...; done; done   # sweep_e.txt
#   python profiles/micro/build_variant.py l64 -DCWTB_MMA_L=64; python profiles/micro/split_timing.py pycwt_b200/variants/libcwtb200_l64.so pycwt_b200/libcwtb200.so   # sweep_q
#   CWTB_EXPAND_MMA=0 python profiles/micro/config_kernels.py 2,4 --prof                   # scalar expansion kernel
#   CWTB_EXPAND_MIN_R=2 python profiles/micro/split_timing.py                              # sweep_t_r4.txt
#   CWTB_BATCH_PIPELINE=0 python bench.py --configs 5                                      # sweep_v
#   python profiles/micro/eps_tradeoff.py; profiles/micro/dmma_vs_dfma; python profiles/micro/e2e_config4.py
"$@"

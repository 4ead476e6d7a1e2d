#!/bin/bash
# partial sums folded into the combine kernel of the staged / end-to-end path: tests, bench, then the final evidence
mkdir -p gpurun_out
bash tools/r2_final_n1.sh

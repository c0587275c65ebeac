#!/bin/bash
# repeats tools/gpu_cfgs.sh timing N times (run-to-run spread).  usage: tools/gpu_cfgs_rep.sh <tag> <configs> <N>
for i in $(seq 1 $3); do bash tools/gpu_cfgs.sh $1 $2; done

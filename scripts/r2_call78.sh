#!/bin/bash
# round-2 combined validation call (2 GPUs): the 1-GPU re-runs of call 8 followed by the 2-GPU tier of call 7
bash scripts/r2_call8.sh
bash scripts/r2_call7.sh

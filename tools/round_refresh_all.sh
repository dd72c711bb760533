#!/bin/bash
R=${1:-r01}
bash tools/round_refresh.sh $R
bash tools/ncu_capture.sh $R

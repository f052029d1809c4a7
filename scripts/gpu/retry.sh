#!/bin/bash
# usage: scripts/gpu/retry.sh <timeout_s> <gpus> <script>   -- retries while the pod answers busy (exit 3 / transient)
T=$1; G=$2; S=$3
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  if [ "$G" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $S" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then echo "[retry $i] busy, sleeping"; sleep 150; else break; fi
done

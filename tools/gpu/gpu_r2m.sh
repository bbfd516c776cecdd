#!/bin/bash
mkdir -p gpurun_out
for pc in 1 4 2 1 4 2; do
  echo "pieces=$pc: $(timeout 300 python tools/query_wall.py 100000 8 12 search_upload_pieces=$pc 2>&1 | tail -1)"
done

#!/bin/bash
# usage: sweep.sh "<opts1>" "<opts2>" ...
for o in "$@"; do
  echo "=== $o"
  python vosk-api_b200/tools/profile_run.py 512 12 "$o" 2 2>&1 | tail -2 | cut -c1-900
done

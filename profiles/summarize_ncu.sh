#!/bin/bash
# usage: summarize_ncu.sh <report.ncu-rep> <out.txt>  -- headline metrics + opcode mix / stall samples of one capture, then drop the report
rep=$1; out=$2
python profiles/ncu_keys.py "$rep" > "$out" 2>&1
ncu -i "$rep" --page source --csv > "${rep%.ncu-rep}_sass.csv" 2>/dev/null
python profiles/sass_summary.py "${rep%.ncu-rep}_sass.csv" >> "$out" 2>&1
rm -f "$rep" "${rep%.ncu-rep}_sass.csv"

#!/bin/bash
# Looks for a Julia toolchain and an installed reference on this machine (the GPU box of a round-end run, a developer's
# workstation); when both are there, dumps the reference's matrices with baseline/run_reference.jl so that
# tests/test_reference_dump.py can compare the engine with the REAL reference instead of the CPU restatement.
# Prints one line of JSON either way.  Nothing here is needed by the product.
cd "$(dirname "$0")/.."
J=$(command -v julia || true)
REF=""
for d in baseline/_ref "$EXTFEM_REFERENCE"; do
  [ -n "$d" ] && [ -f "$d/Project.toml" ] && REF="$d" && break
done
if [ -z "$J" ] || [ -z "$REF" ]; then
  echo "{\"julia\": \"${J:-absent}\", \"reference\": \"${REF:-absent}\", \"dumped\": false}"
  exit 0
fi
mkdir -p baseline/_dump
if timeout 3000 "$J" --project="$REF" -e 'using Pkg; Pkg.instantiate()' >/dev/null 2>&1 && \
   timeout 3000 "$J" --project="$REF" baseline/run_reference.jl baseline/_dump > baseline/_dump/run.log 2>&1; then
  echo "{\"julia\": \"$J\", \"reference\": \"$REF\", \"dumped\": true, \"files\": \"$(ls baseline/_dump/*.bin | tr '\n' ' ')\"}"
else
  echo "{\"julia\": \"$J\", \"reference\": \"$REF\", \"dumped\": false, \"log\": \"baseline/_dump/run.log\"}"
fi

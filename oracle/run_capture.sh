#!/bin/bash
# oracle/run_capture.sh — TEST INFRASTRUCTURE. Runs the reference's own gtest suite (unmodified
# sources, built by `make -C oracle ref_capture`) with Database::Query wrapped by capture_hook.cc,
# leaving capture.jsonl + seg_<n>.bin under oracle/_ref/capture/. tests/golden/make_golden.py
# converts that into the committed golden fixtures.
# Watch.LoadEvents is excluded: it compares unordered_map iteration order and its
# failure path aborts the process.
set -u
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref/capture"
rm -rf "$OUT"
mkdir -p "$OUT"
cd "$HERE/_ref/root/build" || exit 2
VIYA_CAPTURE_DIR="$OUT" "$HERE/_ref/unit_tests_capture" --gtest_filter=-Watch.* > "$OUT/gtest.log" 2>&1
rc=$?
tail -6 "$OUT/gtest.log"
wc -l "$OUT/capture.jsonl"
exit $rc

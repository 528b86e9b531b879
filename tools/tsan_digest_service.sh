#!/bin/bash
# ThreadSanitizer over the HOST side of the library (digest service, operations, cancellation): builds the CPU test
# double (tests/mock_build.py) and the C test program with -fsanitize=thread and runs three concurrent calls with one
# cancel per round.  No GPU needed.  Output -> profiles/r02_tsan_digest_service.txt
set -e
cd "$(dirname "$0")/.."
LIB=$(python -c "from tests import mock_build; print(mock_build.build('thread'))")
W=$(mktemp -d)
gcc -O1 -g -fsanitize=thread -o $W/c3 tests/c/cancel_one_of_three.c -ldl -lpthread
python - "$W" "$LIB" <<'PY'
import hashlib, os, random, subprocess, sys
w, lib = sys.argv[1], sys.argv[2]
rng = random.Random(9)
args = [w + "/c3", lib]
for i in range(3):
    p = f"{w}/f{i}"; d = rng.randbytes(20_000_000); open(p, "wb").write(d); args += [p, hashlib.sha256(d).hexdigest()]
out = subprocess.run(args, capture_output=True, text=True, timeout=900, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0"))
print(out.stdout, end="")
print("ThreadSanitizer reports:", out.stderr.count("WARNING: ThreadSanitizer"))
print(out.stderr[-4000:], end="")
sys.exit(out.returncode)
PY
rm -rf $W
echo "--- mixed load driven from Python (libtsan preloaded; only the library is instrumented)"
LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" python tools/tsan_mixed.py "$LIB" > /tmp/tsan_mixed.$$ 2>&1 || true
tail -1 /tmp/tsan_mixed.$$
echo "ThreadSanitizer reports: $(grep -c 'WARNING: ThreadSanitizer' /tmp/tsan_mixed.$$)"
rm -f /tmp/tsan_mixed.$$

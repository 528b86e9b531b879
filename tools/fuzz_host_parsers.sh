#!/bin/bash
# Mutation fuzzing of the host-side parsers under AddressSanitizer + UBSan (tools/fuzz_host_parsers.py).  No GPU needed.
cd "$(dirname "$0")/.."
export MXD_MOCK_SANITIZE=address,undefined
python -c "from tests import mock_build; mock_build.build('address,undefined')" > /dev/null
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
ASAN_OPTIONS="detect_leaks=0:halt_on_error=0:abort_on_error=0" UBSAN_OPTIONS="print_stacktrace=1" \
  timeout 3000 python tools/fuzz_host_parsers.py > /tmp/fuzz_host.$$ 2>&1
tail -2 /tmp/fuzz_host.$$
echo "AddressSanitizer reports: $(grep -c 'ERROR: AddressSanitizer' /tmp/fuzz_host.$$)"
echo "UBSan reports: $(grep -c 'runtime error:' /tmp/fuzz_host.$$)"
grep -A14 'ERROR: AddressSanitizer\|runtime error:' /tmp/fuzz_host.$$ | head -80
rm -f /tmp/fuzz_host.$$

for hw in 240 120 60 30; do for m in 0 1 2; do echo "== hw=$hw halo=$m"; B200OCR_TC_HALO=$m python tools/bench_conv.py 96 24 3 8 $hw $hw 2>&1 | grep tcconv; done; done

for t in 512 640 768 1024; do WFAGPU_THREADS=$t python tools/cfg5_probe.py 592 | cut -c1-200; done

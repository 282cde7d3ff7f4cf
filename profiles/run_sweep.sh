for i in 1 2; do
timeout 300 python profiles/sweep.py 20 duo:256 duo:296 2>&1 | tail -2
B200FSK_LIB=$PWD/radiosonde_auto_rx_b200/libb200fsk_m2.so timeout 300 python profiles/sweep.py 20 duo:256 duo:296 2>&1 | tail -2
done
B200FSK_LIB=$PWD/radiosonde_auto_rx_b200/libb200fsk_m2p.so timeout 300 python profiles/duo_timing.py 256 262144 2>&1 | tail -8
B200FSK_LIB=$PWD/radiosonde_auto_rx_b200/libb200fsk_m2.so timeout 300 python -m pytest tests/test_fsk_gpu.py -q -m gpu -k "duo" 2>&1 | tail -1

python profiles/sweep.py 18 wpc:2220 wpc:4096 2>&1 | tail -2
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fsk_wpc -s 2 -c 1 python profiles/sweep.py 16 wpc:2220 2>&1 | grep "dram__\|gpu__time" 

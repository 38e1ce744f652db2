for c in A A1 B C; do echo -n "$c: "; python profiles/kernel_times.py $c | tail -1; done

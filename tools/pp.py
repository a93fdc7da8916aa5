import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],2), [round(x,2) for x in d["kernel_ms"]["per_pass_last_step"]], round(d["kernel_ms"]["all_kernels_last_step"],2), 'e2e', d["e2e"]["ms_per_step"] if d.get("e2e") else None)
    except Exception as e: print(f, 'ERR', e)

"""Print the headline fields of gpurun_out/bench_n1.json and the SMO DRAM counters of gpurun_out/smo_dram_r01.csv."""
import csv, json, os
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
try:
    d = json.loads(open(os.path.join(root, "bench_n1.json")).read().strip().splitlines()[-1])
    print("value %.1f fits/s  %.1f ms/step  e2e %.1f  roofline frac %.3f  cpu %.3f  clocks %s  launches %d" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("cpu_baseline", {}).get("value", float("nan")),
        d["clocks"], d["gpu_launches"]))
except Exception as e:
    print("bench_n1.json:", e)
try:
    tot = 0
    for r in csv.reader(open(os.path.join(root, "smo_dram_r01.csv"))):
        if len(r) > 14 and r[0].isdigit():
            print(r[4][:48], r[12], r[14])
            if r[12].startswith("dram__bytes"):
                tot += int(r[14].replace(",", ""))
    print("total DRAM bytes of the SMO launches: %.4e" % tot)
except Exception as e:
    print("smo_dram_r01.csv:", e)

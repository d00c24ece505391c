# dev helper: per-step wall times with different clock samplers (none / nvidia-smi / pynvml)
import sys, time, os, threading, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bcalm_b200.api import Cdbg, MEM_DEVICE
c = Cdbg(k=31, abundance_min=2)
ptr, n = c.synth_reads(5_000_000, 100, 1_500_000, 10000, 42)
def run(tag, steps=12):
    ts = []
    for _ in range(steps):
        torch.cuda.synchronize(0)
        t0 = time.perf_counter(); c.build_raw(ptr, n, MEM_DEVICE, fetch=False); torch.cuda.synchronize(0); ts.append((time.perf_counter() - t0) * 1e3)
    print(tag, " ".join(f"{t:.1f}" for t in ts))
run("warm")
run("none")
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.DEVNULL)
time.sleep(0.5); run("smi clocks only"); p.terminate(); p.wait()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.DEVNULL)
time.sleep(0.5); run("smi clocks+reasons"); p.terminate(); p.wait()
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
stop = False; samples = []
def poll():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.02)
th = threading.Thread(target=poll); th.start(); time.sleep(0.2)
run("pynvml 20ms")
stop = True; th.join()
print("pynvml samples", len(samples), samples[-3:], pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

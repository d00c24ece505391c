"""TEST INFRASTRUCTURE (manual): build the emulated library with ThreadSanitizer (default) or AddressSanitizer (--asan) and run
golden cases on the truly concurrent engine (one OS thread per CUDA thread).  TSan: list the distinct race locations.  ASan: any
out-of-bounds access to a "device" allocation or a static __shared__ array aborts the run (end of round 1: none, for the same
configurations as below plus the sharded path with 3 emulated ranks and the FASTA/FASTQ parser).

    python tests/emu/run_sanitizer.py [--asan | --ubsan] [--sharded] [case ...]     # environment knobs (CDBG_K2C=flat ...) are passed through
(--ubsan: undefined-behaviour sanitizer on the default fiber engine -- shift counts, misaligned vector accesses, signed overflow;
end of round 1: no report over all golden cases.)

Expected, by design (each is a documented lock-free idiom of the kernels, fine under the CUDA memory model used):
  * ld_volatile_u64 on a table key vs. atomicCAS of the same key  (k2_insert, k2c_count, k5_insert, k5c_join, wide/shard tables)
  * plain stores of 1 to an overflow / abandon flag vs. volatile reads of it          (s_over, *overflow)
  * k6_jump / k6_cycle_min: in-place asynchronous pointer jumping on 64-bit words
  * wide_find_or_insert (k >= 32): the tag protocol -- volatile tag reads vs. the CAS that locks a slot, key words published by a fence + tag store
  * kw5s_insert (sharded, k >= 32): every record of a slot stores the same occupancy marker into T.keys[slot]
  * glue_assemble (sharded): neighbouring pieces of a unitig overlap by k-1 characters and both write those (identical) bytes
Anything else -- in particular two plain accesses to shared memory without a barrier in between -- is a bug.  End of round 1: only
the expected locations are reported for the default path, the L2-table / two-pass variants, the opt-in variants, forced splits and
zero-copy input (cases reads_0_k15_a2, lowcx_8_ACGT_k8_a1, circular_test1, even_k_dense_k4)."""
import glob
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import build_emu  # noqa: E402


def main():
    asan = "--asan" in sys.argv[1:]
    ubsan = "--ubsan" in sys.argv[1:]
    sharded = "--sharded" in sys.argv[1:]          # 3 emulated ranks (host threads + the NCCL stand-in), with -all-abundance-counts
    cases = [a for a in sys.argv[1:] if a not in ("--asan", "--ubsan", "--sharded")] or ["reads_0_k15_a2", "lowcx_8_ACGT_k8_a1", "circular_test1", "even_k_dense_k4"]
    build_emu.build()
    work = tempfile.mkdtemp(prefix="cdbg_tsan_")
    src = os.path.join(work, "a", "b")
    os.makedirs(src)
    os.makedirs(os.path.join(work, "include"))
    for f in os.listdir(build_emu.OUT):
        if f.endswith((".cu", ".cuh", ".inl", ".h")):
            with open(os.path.join(build_emu.OUT, f)) as fi, open(os.path.join(src, f), "w") as fo:
                fo.write(fi.read())
    with open(os.path.join(ROOT, "include", "cdbg.h")) as fi, open(os.path.join(work, "include", "cdbg.h"), "w") as fo:
        fo.write(fi.read())
    lib = os.path.join(work, "libcdbg_tsan.so")
    subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-fno-strict-aliasing", "-w",
                           "-fsanitize=address" if asan else "-fsanitize=undefined" if ubsan else "-fsanitize=thread", "-fno-omit-frame-pointer", "-I", HERE, "-o", lib,
                           os.path.join(src, "cdbg.cu"), "-ldl"])
    tsan = subprocess.check_output(["gcc", "-print-file-name=" + ("libasan.so" if asan else "libubsan.so" if ubsan else "libtsan.so")], text=True).strip()
    code = (
        "import sys, json\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from bcalm_b200 import api, parity\n"
        f"api.LIB_PATH = {lib!r}\n"
        f"cases = {{c['name']: c for c in json.load(open({os.path.join(ROOT, 'tests', 'golden', 'golden.json')!r}))}}\n"
        f"sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\n"
        f"sharded = {sharded!r}\n"
        "for n in sys.argv[1:]:\n"
        "    c = cases[n]\n"
        "    if sharded:\n"
        "        import test_emu_sharded as tes\n"
        "        g, _ = tes.sharded_build(c['reads'], c['k'], c['abundance_min'], 3, all_abundance_counts=True)\n"
        "    else:\n"
        "        with api.Cdbg(k=c['k'], abundance_min=c['abundance_min']) as x:\n"
        "            g = x.build(c['reads'])\n"
        "    print(n, parity.diff(parity.graph_to_records(g), parity.parse_unitigs_fa(c['unitigs_fa']), c['k']) or 'OK', flush=True)\n")
    env = dict(os.environ, CDBG_EMU_ENGINE="fibers" if ubsan else "threads", LD_PRELOAD=tsan,
               TSAN_OPTIONS=f"halt_on_error=0 log_path={os.path.join(work, 'report')} history_size=3",
               ASAN_OPTIONS="detect_leaks=0:halt_on_error=1")
    rc = subprocess.run([sys.executable, "-c", code, *cases], env=env).returncode
    if asan or ubsan:
        print("sanitizer run:", "finished" if rc == 0 else f"FAILED (exit {rc})", "(UBSan prints its findings above)" if ubsan else "")
        sys.exit(rc)
    locs = {}
    for rep in glob.glob(os.path.join(work, "report.*")):
        text = open(rep).read()
        for block in text.split("WARNING: ThreadSanitizer")[1:]:
            frames = re.findall(r"#\d+ (.+?) \S*a/b/(\S+)", block)
            if frames:
                fn, where = frames[0] if "ld_volatile" not in frames[0][0] or len(frames) < 2 else frames[1]
                key = f"{where}  {fn.split('(')[0]}"
                locs[key] = locs.get(key, 0) + 1
    print("distinct race locations:")
    for k, v in sorted(locs.items()):
        print(f"  {v:3d}  {k}")


if __name__ == "__main__":
    main()

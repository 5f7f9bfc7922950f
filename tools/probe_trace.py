import gzip, os, sys, time, torch
sys.path.insert(0, "/root/repo")
ROOT="/root/repo"
from swiftover_b200 import ChainFile, synth as S
dev = torch.device("cuda", 0)
chain = gzip.open(os.path.join(ROOT, "tests/golden/resources/hg19ToHg38.over.chain.gz"), "rb").read()
cf = ChainFile(text=chain)
N = 100_000_000
iv = S.gen_intervals(cf, N, 20240002, device=dev, sort=True)
text = S.bed_text(cf, iv, ncols=3)
host = torch.empty(text.numel(), dtype=torch.uint8, pin_memory=True); host.copy_(text); torch.cuda.synchronize()
del iv, text; torch.cuda.empty_cache()
for _ in range(3):
    t0 = time.perf_counter(); cf.lift_bed_text_ptr(host.data_ptr(), host.numel()); print("ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr)

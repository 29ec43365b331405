import sys, torch
sys.path.insert(0, '/root/repo')
from nnx_ppo_b200 import _lib
lib = _lib.load()
src = torch.randn(1 << 20, device='cuda')
out = torch.zeros(4, dtype=torch.int64, device='cuda')
for blocks in (1, 148):
    for N in (16, 64, 128, 256):
        for byts in (4096, 32768, 65536):
            _lib.check(lib.b200ppo_tc_microbench(_lib.current_stream(), src.data_ptr(), out.data_ptr(), N, 512, byts, min(4, 512 // N), blocks))
            torch.cuda.synchronize()
            o = out.cpu().tolist()
            print(f"blocks={blocks} N={N} bytes={byts}: cyc/mma dep={o[0]/512:.1f} multi-acc={o[2]/512:.1f} | cyc/copy serial={o[1]/512:.0f} ({byts/(o[1]/512):.1f} B/clk) 4-split={o[3]/128:.0f} ({byts/(o[3]/128):.1f} B/clk)")

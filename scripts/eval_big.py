import sys, time, torch
sys.path.insert(0, '/root/repo')
from ibinn_b200 import config
from ibinn_b200.model import GenerativeClassifier
dev = torch.device('cuda:0')
torch.manual_seed(0)
m = GenerativeClassifier(config.make_args()).to(dev)
m.eval()
for B in (500, 2048):
    x = torch.randn(B, 3, 32, 32, device=dev)
    y = torch.zeros(B, 10, device=dev); y[torch.arange(B), torch.randint(0, 10, (B,))] = 1.
    with torch.no_grad():
        out = m.validate(x, y)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): out = m.validate(x, y)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(B, {k: float(v.mean()) for k, v in out.items() if torch.is_tensor(v) and v.numel() >= 1 and k != 'logits_val'}, '%.1f ms  %.0f img/s' % (dt * 1e3, B / dt))

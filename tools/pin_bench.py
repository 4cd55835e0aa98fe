import time, torch
torch.cuda.init(); x=torch.zeros(1,device='cuda')
for name, f in [("tensor.pin_memory", lambda: torch.tensor(list(range(64)), dtype=torch.int32).pin_memory()),
                ("empty(pin_memory=True)", lambda: torch.empty(65, dtype=torch.int32, pin_memory=True)),
                ("tensor(device=cuda)", lambda: torch.tensor(list(range(64)), dtype=torch.int32, device='cuda')),
                ("pin+to", lambda: torch.tensor(list(range(64)), dtype=torch.int32).pin_memory().to('cuda', non_blocking=True))]:
    for rep in range(3):
        torch.cuda.synchronize(); t=time.perf_counter()
        for _ in range(50): y=f()
        torch.cuda.synchronize(); print(name, rep, f"{(time.perf_counter()-t)/50*1e6:.1f} us")

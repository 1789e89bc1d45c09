"""CPU: host-side logic that needs no GPU."""
import torch


def test_brdf_phase_lr_matches_guarded_steplr():
    from materialist_b200.inverse import brdf_phase_lr
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=3e-4)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=100, gamma=0.8)
    for k in range(650):
        lr = opt.param_groups[0]["lr"]
        assert abs(lr - brdf_phase_lr(k)) < 1e-12, k
        opt.step()
        if lr > 1.5e-4:
            sched.step()
    assert abs(brdf_phase_lr(10_000) - 3e-4 * 0.8 ** 4) < 1e-12

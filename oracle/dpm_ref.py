"""Oracle: diffusion glue + DPM-Solver++ sampling loop (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, for the configuration every shipped SlotDiffusion config uses
(eps-prediction, linear beta schedule, DPM-Solver++ singlestep order 3,
20 NFE, time_uniform, vq_denoised, no guidance):
  * make_beta_schedule 'linear'        video_based/models/ddpm/utils.py:21-27
  * DDPM.register_schedule buffers     video_based/models/ddpm/ddpm.py:69-131
  * q_sample (_sample_xt_from_x0)      ddpm.py:161-165
  * LDM.loss_function (eps target)     video_based/models/ddpm/ldm.py:59-83
  * NoiseScheduleVP('discrete')        video_based/models/ddpm/dpm_solver.py:160-235 (+ interpolate_fn :11-50)
  * model_wrapper time map             dpm_solver.py:339-348
  * data_prediction_fn + vq_denoised   dpm_solver.py:523-534
  * singlestep order plan / 2nd / 3rd order updates   dpm_solver.py:574-631, :690-735, :767-831
  * sample() singlestep loop           dpm_solver.py:1310-1328
  * CondDDPM.generate_imgs DPM branch  video_based/models/ddpm/cond_ddpm.py:155-193
  * VectorQuantizer nearest-code       video_based/models/vqvae/quantize.py:80-93
(all under /root/reference/slotdiffusion/).  Scalars follow the reference's
fp32 tensor arithmetic so that coefficients agree to the last bits.
"""
import numpy as np
import torch


def linear_betas(linear_start=0.0015, linear_end=0.0195, n=1000):
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n, dtype=torch.float64) ** 2).numpy()


def ddpm_buffers(betas):
    ab = np.cumprod(1.0 - betas, axis=0)
    f = lambda a: torch.tensor(a, dtype=torch.float32)
    return dict(betas=f(betas), alphas_bar=f(ab), sqrt_alphas_bar=f(np.sqrt(ab)),
                sqrt_one_minus_alphas_bar=f(np.sqrt(1.0 - ab)))


def q_sample(buf, x0, t, noise):
    a = buf['sqrt_alphas_bar'][t].to(x0.dtype).view(-1, 1, 1, 1)
    s = buf['sqrt_one_minus_alphas_bar'][t].to(x0.dtype).view(-1, 1, 1, 1)
    return a * x0 + s * noise


def denoise_loss(unet_fn, buf, x0, t, noise, context):
    """mean((eps_theta(x_t, t, ctx) - eps)^2), ldm.py:66-82."""
    pred = unet_fn(q_sample(buf, x0, t, noise), t, context)
    return ((pred - noise) ** 2).mean()


def vq_quantize(z, codebook):
    """z [B,C,h,w], codebook [K,C] -> nearest code per pixel (quantize.py:84-94).
    Distance uses the reference's expanded form ||z||^2 + ||e||^2 - 2 z.e."""
    B, C, H, W = z.shape
    zf = z.permute(0, 2, 3, 1).reshape(-1, C)
    e = codebook.to(z.dtype)
    d = (zf ** 2).sum(1, keepdim=True) + (e ** 2).sum(1) - 2 * zf @ e.t()
    idx = torch.argmin(d, dim=1)
    zq = e[idx].view(B, H, W, C).permute(0, 3, 1, 2).contiguous()
    return zq, idx.view(B, H, W)


class NoiseScheduleVP:
    """Discrete-time schedule, dpm_solver.py:160-235."""

    def __init__(self, betas_f32):
        self.log_alpha = (0.5 * torch.log(1 - betas_f32).cumsum(dim=0)).float()
        self.N = len(self.log_alpha)
        self.t_arr = torch.linspace(0., 1., self.N + 1)[1:].float()

    @staticmethod
    def _interp(x, xp, yp):
        """Piece-wise linear through (xp, yp) (xp ascending), linear extrapolation with
        the outermost segments -- the behaviour of interpolate_fn (dpm_solver.py:11-50)."""
        K = xp.numel()
        # number of keypoints strictly below x (ties: x sorts first, as torch.sort is stable
        # and x is concatenated in front, dpm_solver.py:25-27)
        pos = torch.searchsorted(xp, x, right=False)
        lo = torch.clamp(pos - 1, 0, K - 2)
        x0, x1, y0, y1 = xp[lo], xp[lo + 1], yp[lo], yp[lo + 1]
        return y0 + (x - x0) * (y1 - y0) / (x1 - x0)

    def log_mean_coeff(self, t):
        return self._interp(t, self.t_arr, self.log_alpha)

    def alpha(self, t):
        return torch.exp(self.log_mean_coeff(t))

    def std(self, t):
        return torch.sqrt(1. - torch.exp(2. * self.log_mean_coeff(t)))

    def lam(self, t):
        la = self.log_mean_coeff(t)
        return la - 0.5 * torch.log(1. - torch.exp(2. * la))

    def inverse_lambda(self, lamb):
        la = -0.5 * torch.logaddexp(torch.zeros(1), -2. * lamb)
        return self._interp(la, torch.flip(self.log_alpha, [0]), torch.flip(self.t_arr, [0]))


def singlestep_plan(steps=20, order=3):
    """orders + outer time grid, dpm_solver.py:574-631 (time_uniform)."""
    assert order == 3
    K = steps // 3 + 1
    if steps % 3 == 0:
        orders = [3] * (K - 2) + [2, 1]
    elif steps % 3 == 1:
        orders = [3] * (K - 1) + [1]
    else:
        orders = [3] * (K - 1) + [2]
    return orders


def dpm_coefficients(ns, steps=20, order=3):
    """All per-evaluation scalars of the 20-NFE run, as python floats.

    Returns a list of outer steps; each: dict(order, evals=[dict(t_model, alpha, sigma)...],
    and the linear-combination weights of the updates (dpm_solver.py:716-732, :804-831)).
    """
    orders = singlestep_plan(steps, order)
    t_T, t_0 = 1.0, 1.0 / ns.N
    grid = torch.linspace(t_T, t_0, steps + 1)
    outer = grid[torch.cumsum(torch.tensor([0] + orders), 0)]
    plan = []
    for i, o in enumerate(orders):
        s, t = outer[i], outer[i + 1]
        inner = torch.linspace(s.item(), t.item(), o + 1)
        lam_in = ns.lam(inner)
        h_in = lam_in[-1] - lam_in[0]
        r1 = None if o <= 1 else (lam_in[1] - lam_in[0]) / h_in
        r2 = None if o <= 2 else (lam_in[2] - lam_in[0]) / h_in
        s1v, t1v = s.reshape(1), t.reshape(1)
        lam_s, lam_t = ns.lam(s1v), ns.lam(t1v)
        h = lam_t - lam_s
        st = dict(order=o)

        def ev(tt):
            return dict(t_cont=tt, t_model=(tt - 1. / ns.N) * 1000., alpha=ns.alpha(tt), sigma=ns.std(tt))
        if o == 1:
            st['evals'] = [ev(s1v)]
            st['x_t'] = dict(x=ns.std(t1v) / ns.std(s1v), m_s=-(ns.alpha(t1v) * torch.expm1(-h)))
        elif o == 2:
            s1 = ns.inverse_lambda(lam_s + r1 * h)
            st['evals'] = [ev(s1v), ev(s1)]
            phi_11, phi_1 = torch.expm1(-r1 * h), torch.expm1(-h)
            st['x_s1'] = dict(x=ns.std(s1) / ns.std(s1v), m_s=-(ns.alpha(s1) * phi_11))
            st['x_t'] = dict(x=ns.std(t1v) / ns.std(s1v), m_s=-(ns.alpha(t1v) * phi_1),
                             d1=-(0.5 / r1) * (ns.alpha(t1v) * phi_1))
        else:
            s1 = ns.inverse_lambda(lam_s + r1 * h)
            s2 = ns.inverse_lambda(lam_s + r2 * h)
            st['evals'] = [ev(s1v), ev(s1), ev(s2)]
            phi_11, phi_12, phi_1 = torch.expm1(-r1 * h), torch.expm1(-r2 * h), torch.expm1(-h)
            phi_22 = torch.expm1(-r2 * h) / (r2 * h) + 1.
            phi_2 = phi_1 / h + 1.
            sig_s = ns.std(s1v)
            st['x_s1'] = dict(x=ns.std(s1) / sig_s, m_s=-(ns.alpha(s1) * phi_11))
            st['x_s2'] = dict(x=ns.std(s2) / sig_s, m_s=-(ns.alpha(s2) * phi_12),
                              d1=r2 / r1 * (ns.alpha(s2) * phi_22))
            st['x_t'] = dict(x=ns.std(t1v) / sig_s, m_s=-(ns.alpha(t1v) * phi_1),
                             d2=(1. / r2) * (ns.alpha(t1v) * phi_2))
        plan.append(st)
    return plan


def dpm_sample(unet_fn, betas_f32, x_T, context, codebook=None, steps=20, return_trace=False):
    """20-NFE DPM-Solver++ singlestep order-3 sampling of latents.

    unet_fn(x, t_model[B] float, context) -> eps.  codebook [K,C] enables vq_denoised
    (LDM default, ldm.py:56-57); None disables it.
    """
    ns = NoiseScheduleVP(betas_f32)
    plan = dpm_coefficients(ns, steps)
    B = x_T.shape[0]
    trace = []

    def m(x, e):
        eps = unet_fn(x, e['t_model'].expand(B).float(), context)
        x0 = (x - e['sigma'].to(x.dtype) * eps) / e['alpha'].to(x.dtype)      # dpm_solver.py:529-530
        if codebook is not None:
            x0 = vq_quantize(x0, codebook)[0]                                 # :532-533
        if return_trace:
            trace.append(dict(x=x, eps=eps, x0=x0))
        return x0
    x = x_T
    for st in plan:
        ev = st['evals']
        m_s = m(x, ev[0])
        c = {k: {kk: vv.to(x.dtype) for kk, vv in st[k].items()} for k in st if k.startswith('x_')}
        if st['order'] == 1:
            x = c['x_t']['x'] * x + c['x_t']['m_s'] * m_s
        elif st['order'] == 2:
            x_s1 = c['x_s1']['x'] * x + c['x_s1']['m_s'] * m_s
            m_s1 = m(x_s1, ev[1])
            x = c['x_t']['x'] * x + c['x_t']['m_s'] * m_s + c['x_t']['d1'] * (m_s1 - m_s)
        else:
            x_s1 = c['x_s1']['x'] * x + c['x_s1']['m_s'] * m_s
            m_s1 = m(x_s1, ev[1])
            x_s2 = c['x_s2']['x'] * x + c['x_s2']['m_s'] * m_s + c['x_s2']['d1'] * (m_s1 - m_s)
            m_s2 = m(x_s2, ev[2])
            x = c['x_t']['x'] * x + c['x_t']['m_s'] * m_s + c['x_t']['d2'] * (m_s2 - m_s)
    if return_trace:
        return x, trace
    return x

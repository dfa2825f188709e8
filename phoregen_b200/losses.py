"""Value of the training objective, `PhoreDiff.compute_loss` (models/diffusion.py:249-352), forward only.

What the reference's validation loop needs (compute_loss under no_grad): the noise levels, the forward-diffused inputs,
one forward pass (on the CUDA kernels) and the loss terms.  The backward pass is not built, so the returned loss carries
no autograd graph; training with this package is a later tier.

Every random draw is made in the reference's order (sample_time -> position noise -> node Gumbel noise -> edge Gumbel
noise) with the global torch generator on `rng_device`, so equal seeds give the reference's draws when both run on the
same device kind (tests/test_cpu_losses.py compares on the CPU).
"""
import torch
import torch.nn.functional as F

EPS = 1e-30


def sample_time(num_graphs, num_timesteps, device):
    """diffusion.py:138-145: antithetic pairs t, T-1-t."""
    t = torch.randint(0, num_timesteps, size=(num_graphs // 2 + 1,), device=device)
    return torch.cat([t, num_timesteps - t - 1], dim=0)[:num_graphs]


def index_to_log_onehot(x, num_classes):
    return torch.log(F.one_hot(x, num_classes).float().clamp(min=1e-30))            # common.py:398-402


def gaussian_add_noise(alphas_bar, x, time_step, batch):
    """ContigousTransition.add_noise, continuous branch (transition.py:28-41)."""
    a_bar = alphas_bar.index_select(0, time_step).index_select(0, batch).unsqueeze(-1)
    noise = torch.zeros_like(x).normal_()
    return a_bar.sqrt() * x + (1 - a_bar).sqrt() * noise


def categorical_add_noise(q_mats, v, time_step, batch, num_classes):
    """GeneralCategoricalTransition.add_noise (transition.py:244-267): sample v_t ~ q(v_t | v_0) with Gumbel noise.
    -> (one-hot v_t, log one-hot v_t, log one-hot v_0)."""
    log_v0 = index_to_log_onehot(v, num_classes)
    qt = q_mats[time_step][batch]
    log_q = torch.log(torch.einsum("...i,...ij->...j", log_v0.exp(), qt) + EPS).clamp_min(-32.0)
    uniform = torch.rand_like(log_q)                                                 # common.py:425-431
    cls = (-torch.log(-torch.log(uniform + 1e-30) + 1e-30) + log_q).argmax(dim=-1)
    return F.one_hot(cls, num_classes).float(), index_to_log_onehot(cls, num_classes), log_v0


def q_v_posterior(q_mats, q_onestep_t, log_v0, log_vt, t, batch):
    """transition.py:285-315 with v0_prob=True."""
    tm1 = torch.where(t - 1 < 0, torch.zeros_like(t), t - 1)
    fact1 = torch.einsum("bj,bjk->bk", torch.exp(log_vt), q_onestep_t[t][batch])
    fact2 = torch.einsum("bj,bjk->bk", torch.exp(log_v0), q_mats[tm1][batch])
    out = torch.log(fact1 + EPS).clamp_min(-32.0) + torch.log(fact2 + EPS).clamp_min(-32.0)
    out = out - torch.logsumexp(out, dim=-1, keepdim=True)
    return torch.where(t[batch].unsqueeze(-1) == 0, log_v0, out)


def v_Lt(log_post_true, log_post_pred, log_v0, t, batch):
    """transition.py:317-329: KL between the posteriors, decoder NLL at t = 0."""
    kl = (log_post_true.exp() * (log_post_true - log_post_pred)).sum(dim=-1)
    nll = -(log_v0.exp() * log_post_pred).sum(dim=-1)
    mask = (t == 0).float()[batch]
    return mask * nll + (1 - mask) * kl


def qd_loss(y_true, y_l, y_u, a=0.05, s=160, nd=15, factor=1, epsilon=1e-12):
    """Quality-driven interval loss, soft mode (common.py:261-281)."""
    n = y_true.shape[0]
    k_h = torch.relu(torch.sign(y_u - y_true)) * torch.relu(torch.sign(y_true - y_l))
    k_s = torch.sigmoid((y_u - y_true) * s) * torch.sigmoid((y_true - y_l) * s)
    mpiw = torch.sum((y_u - y_l) * k_h) / (torch.sum(k_h) + epsilon) * factor
    return mpiw + (torch.relu((1 - a) - torch.mean(k_s)) ** 2) * (n ** 0.5) * nd


def _exact_match_fraction(true_cls, logits, batch, n_graphs):
    """common.py:284-297: fraction of molecules whose arg-max classes are all right."""
    wrong = (logits.argmax(dim=-1) != true_cls).float()
    per_mol = torch.zeros(n_graphs, device=wrong.device).index_add_(0, batch, wrong)
    present = torch.zeros(n_graphs, device=wrong.device).index_add_(0, batch, torch.ones_like(wrong)) > 0
    return int(((per_mol == 0) & present).sum().item()) / int(present.sum().item())


def perturb(model, pos, x, batch_node, f_edge_attr, batch_edge, n_graphs, rng_device=None):
    """Steps 1-2 of compute_loss (diffusion.py:250-265).  Draws happen on `rng_device` (default: the data's device)."""
    dev = pos.device
    rd = torch.device(rng_device) if rng_device is not None else dev
    to = lambda t: t.to(rd)
    t = sample_time(n_graphs, model.num_timesteps, rd)
    pos_pert = gaussian_add_noise(to(model.pos_transition.alphas_bar), to(pos), t, to(batch_node))
    h_node, log_node_t, log_node_0 = categorical_add_noise(to(model.node_transition.q_mats), to(x), t, to(batch_node), model.num_node_types)
    h_edge, log_edge_t, log_edge_0 = categorical_add_noise(to(model.edge_transition.q_mats), to(f_edge_attr), t, to(batch_edge), model.num_edge_types)
    out = dict(time_step=t, pos_pert=pos_pert, h_node_pert=h_node, log_node_t=log_node_t, log_node_0=log_node_0,
               h_edge_pert=h_edge, log_edge_t=log_edge_t, log_edge_0=log_edge_0)
    return {k: v.to(dev) for k, v in out.items()}


def loss_terms(model, pert, preds, pos, x, batch_node, f_edge_attr, batch_edge, num_atoms, bond_edge_index=None):
    """Step 4 of compute_loss (diffusion.py:281-351) -> (loss_total, loss_dict)."""
    pred_node, pred_pos, pred_edge, pred_count = preds
    t = pert["time_step"]
    n_graphs = int(t.numel())
    w = model.loss_weight
    loss_pos = F.mse_loss(pred_pos, pos) * w[0]
    nt, et = model.node_transition, model.edge_transition
    log_node_recon = F.log_softmax(pred_node, dim=-1)
    post_true = q_v_posterior(nt.q_mats, nt.transpopse_q_onestep_mats, pert["log_node_0"], pert["log_node_t"], t, batch_node)
    post_pred = q_v_posterior(nt.q_mats, nt.transpopse_q_onestep_mats, log_node_recon, pert["log_node_t"], t, batch_node)
    loss_node = torch.mean(v_Lt(post_true, post_pred, pert["log_node_0"], t, batch_node)) * w[1]
    log_edge_recon = F.log_softmax(pred_edge, dim=-1)
    post_true = q_v_posterior(et.q_mats, et.transpopse_q_onestep_mats, pert["log_edge_0"], pert["log_edge_t"], t, batch_edge)
    post_pred = q_v_posterior(et.q_mats, et.transpopse_q_onestep_mats, log_edge_recon, pert["log_edge_t"], t, batch_edge)
    loss_edge = torch.mean(v_Lt(post_true, post_pred, pert["log_edge_0"], t, batch_edge)) * w[2]
    true_count = (num_atoms.to(pred_pos.device) - model.min_atom) / (model.max_atom - model.min_atom)
    loss_count = qd_loss(true_count.unsqueeze(-1).float(), *pred_count, s=160, nd=15, factor=model.count_factor)
    total = loss_pos + loss_node + loss_edge + loss_count
    d = {"loss_pos": loss_pos.item(), "loss_node": loss_node.item(), "loss_count": loss_count.item()}
    if model.bond_len_loss:
        src, dst = bond_edge_index
        loss_len = F.mse_loss(torch.norm(pred_pos[src] - pred_pos[dst], dim=-1), torch.norm(pos[src] - pos[dst], dim=-1))
        total = total + loss_len
        d["loss_len"] = loss_len.item()
    d["loss_edge"] = loss_edge.item()
    d["loss"] = total.item()
    d["node_acc"] = _exact_match_fraction(x, pred_node, batch_node, n_graphs)
    d["edge_acc"] = _exact_match_fraction(f_edge_attr, pred_edge, batch_edge, n_graphs)
    return total, d


def compute_loss(model, data, forward=None, rng_device=None):
    """`PhoreDiff.compute_loss(data)` for a PyG-style batch (data['ligand'].{x,pos,batch,ptr}, data['ligand','ligand'].
    {f_edge_attr,f_edge_index,f_edge_attr_batch[,edge_index]}, data['phore'].{x,pos,norm,batch}, data.num_graphs).
    `forward` defaults to the model's CUDA forward; tests inject the reference's predictions."""
    lig, ll, ph = data["ligand"], data["ligand", "ligand"], data["phore"]
    G = int(data.num_graphs)
    pert = perturb(model, lig.pos, lig.x, lig.batch, ll.f_edge_attr, ll.f_edge_attr_batch, G, rng_device)
    fwd = forward or model.forward
    preds = fwd(h_node_pert=pert["h_node_pert"], pos_pert=pert["pos_pert"], batch_node=lig.batch, h_edge_pert=pert["h_edge_pert"],
                edge_index=ll.f_edge_index, batch_edge=ll.f_edge_attr_batch, time_step=pert["time_step"], h_phore=ph.x,
                pos_phore=ph.pos, phore_norm=ph.norm, batch_phore=ph.batch)
    num_atoms = lig.ptr[1:] - lig.ptr[:-1]
    return loss_terms(model, pert, preds, lig.pos, lig.x, lig.batch, ll.f_edge_attr, ll.f_edge_attr_batch, num_atoms,
                      bond_edge_index=getattr(ll, "edge_index", None))

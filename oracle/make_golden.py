#!/usr/bin/env python3
"""Generate tests/golden/ref_shim_*.npz by executing the reference's OWN source files
(/root/reference/signals.py, /root/reference/model.py, unmodified, imported from where
they lie) over the torch-backed TensorFlow shim in oracle/tf_shim.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference checkout does not
exist on the GPU box; the committed .npz fixtures are what travels):

    python oracle/make_golden.py [--ref /root/reference] [--out tests/golden]

Also writes tests/golden/kat_appendix_b.npz (SURVEY.md Appendix B known answers).
"""
import argparse
import configparser
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    ap.add_argument('--out', default=os.path.join(os.path.dirname(HERE), 'tests', 'golden'))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)

    sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
    sys.path.insert(1, args.ref)
    import tensorflow as tf          # the shim
    import torch
    import signals as ref_signals    # the reference, unmodified
    import model as ref_model        # the reference, unmodified

    cfg = configparser.ConfigParser()
    cfg.read(os.path.join(args.ref, 'config'))
    params = cfg['DEFAULT']
    params['simulate_noise'] = 'False'          # as train.py:256 does before building the fine-tuner

    rng = np.random.default_rng(20261018)

    # ------------------------------------------------------------------ A. forward + tape.gradient
    n = 192
    oef = rng.uniform(0.04, 0.84, n)
    dbv = rng.uniform(0.001, 0.201, n)
    x = np.stack([oef, dbv], -1).astype(np.float32)
    x[:4] = [[0.4, 0.12], [0.4, 0.03], [0.04, 0.001], [0.84, 0.201]]
    g_rand = rng.standard_normal((n, 11)).astype(np.float32)
    out = {'oef_dbv': x, 'g_rand': g_rand}
    for full in (True, False):
        for blood in (True, False):
            layer = ref_signals.SignalGenerationLayer(params, full, blood)
            key = 'f%d_b%d' % (int(full), int(blood))
            inp = tf.convert_to_tensor(x.reshape(n, 1, 1, 1, 2))
            with tf.GradientTape(persistent=True) as tape:
                tape.watch(inp)
                o = layer(inp)
                o_w = o * tf.convert_to_tensor(g_rand.reshape(n, 1, 1, 1, 11))
            out['signal_' + key] = o.detach().numpy().reshape(n, 11)
            out['grad_ones_' + key] = tape.gradient(o, inp).numpy().reshape(n, 2)
            out['grad_rand_' + key] = tape.gradient(o_w, inp).numpy().reshape(n, 2)
            out['taus'] = layer._taus.numpy()
    np.savez(os.path.join(args.out, 'ref_shim_forward.npz'), **out)
    print('forward: demo input (signals.py:309) ->', out['signal_f1_b1'][0])

    # ------------------------------------------------------------------ B. ELBO pieces
    def run_elbo(tag, multi_norm, df, seed):
        B, X, Y, Z = 2, 4, 4, 2
        nv = B * X * Y * Z
        r = np.random.default_rng(seed)
        q = np.stack([r.normal(-0.3, 0.7, nv), r.normal(0.0, 0.6, nv), r.normal(-1.2, 0.7, nv),
                      r.normal(0.0, 0.6, nv), r.normal(0.0, 0.8, nv)], -1).astype(np.float32)
        prior = (q + r.normal(0, 0.3, (nv, 5))).astype(np.float32)
        sigma = np.exp(r.normal(np.log(0.05), 0.2, (nv, 11))).astype(np.float32)
        mask = (r.uniform(size=nv) > 0.25).astype(np.float32)
        truth = np.stack([r.uniform(0.1, 0.7, nv), r.uniform(0.005, 0.15, nv)], -1).astype(np.float32)
        sig_layer = ref_signals.SignalGenerationLayer(params, True, True)
        data = sig_layer(tf.convert_to_tensor(truth)).numpy() * 100.0
        data = (data * (1.0 + 0.02 * r.standard_normal((nv, 11)))).astype(np.float32)
        data = data * mask[:, None]                                   # train.py:56 pre-masks the data

        trainer = ref_model.EncoderTrainer(system_params=params, no_units=60, no_intermediate_layers=2,
                                           student_t_df=df, initial_im_sigma=0.05,
                                           multi_image_normalisation=multi_norm, channelwise_gating=True,
                                           infer_inv_gamma=False, use_population_prior=False, use_mvg=True,
                                           predict_log_data=False)
        shp = (B, X, Y, Z)
        q_t = tf.convert_to_tensor(q.reshape(shp + (5,)))
        s_t = tf.convert_to_tensor(sigma.reshape(shp + (11,)))
        m_t = tf.convert_to_tensor(mask.reshape(shp + (1,)))
        d_t = tf.convert_to_tensor(data.reshape(shp + (11,)))
        p_t = tf.convert_to_tensor(prior.reshape(shp + (5,)))
        q_t.requires_grad_(True)
        s_t.requires_grad_(True)
        tf.random.set_seed(seed)
        tf.random.LOG.clear()
        sampled = ref_model.ReparamTrickLayer(trainer)((q_t, m_t))     # model.py:248
        pred = sig_layer(sampled)                                      # model.py:273
        y_pred = tf.concat([pred, s_t], -1)                            # model.py:276
        y_true = tf.concat([d_t, m_t], -1)                             # train.py:58,62
        nll = trainer.fine_tune_loss_fn(y_true, y_pred)
        nll_map = trainer.fine_tune_loss_fn(y_true, y_pred, return_mean=False)
        n_before_kl = len(tf.random.LOG)
        kl = trainer.kl_loss(tf.concat([p_t, m_t], -1), q_t)           # 70 samples (model.py:654)
        eps = tf.random.LOG[0][1].numpy().reshape(nv, 2)
        eps_kl = np.stack([e[1].numpy().reshape(nv, 2) for e in tf.random.LOG[n_before_kl:]], 1)
        assert eps_kl.shape == (nv, 70, 2), eps_kl.shape
        g_nll = torch.autograd.grad(nll, [q_t, s_t], retain_graph=True)
        g_kl = torch.autograd.grad(kl, [q_t], retain_graph=True)
        # second, independent KL evaluation for the per-voxel map (fresh draws are recorded too)
        n0 = len(tf.random.LOG)
        kl_map = trainer.kl_loss(tf.concat([p_t, m_t], -1), q_t, return_mean=False, no_samples=70)
        eps_kl2 = np.stack([e[1].numpy().reshape(nv, 2) for e in tf.random.LOG[n0:]], 1)
        np.savez(os.path.join(args.out, 'ref_shim_elbo_%s.npz' % tag),
                 q=q, prior=prior, sigma=sigma, mask=mask, data=data, eps=eps, eps_kl=eps_kl,
                 sampled=sampled.detach().numpy().reshape(nv, 2), pred=pred.detach().numpy().reshape(nv, 11),
                 nll=nll.detach().numpy(), kl=kl.detach().numpy(),
                 nll_map=nll_map.detach().numpy().reshape(nv),
                 grad_q_nll=g_nll[0].numpy().reshape(nv, 5), grad_sigma=g_nll[1].numpy().reshape(nv, 11),
                 grad_q_kl=g_kl[0].numpy().reshape(nv, 5),
                 eps_kl2=eps_kl2, kl_map2=kl_map.detach().numpy().reshape(nv),
                 multi_image_normalisation=multi_norm, student_t_df=df, se_idx=trainer._se_idx)
        print('elbo[%s]: nll=%.6f kl=%.6f' % (tag, float(nll), float(kl)))
        return trainer, q

    trainer, q = run_elbo('optimal', False, 200, 11)
    run_elbo('multinorm', True, 200, 12)
    run_elbo('studentt', False, 10, 13)

    # ------------------------------------------------------------------ C. calculate_means
    nv = q.shape[0]
    tf.random.set_seed(5)
    tf.random.LOG.clear()
    q5 = tf.convert_to_tensor(q.reshape(2, 4, 4, 2, 5))
    means, stds = trainer.calculate_means(q5, tf.ones_like(q5[:, :, :, :, :1]), include_r2p=True,
                                          return_stds=True, no_samples=16)
    eps_s = np.stack([e[1].numpy().reshape(nv, 2) for e in tf.random.LOG], 1)
    np.savez(os.path.join(args.out, 'ref_shim_means.npz'), q=q, eps=eps_s,
             means=means.numpy().reshape(nv, 3), stds=stds.numpy().reshape(nv, 3))

    # ------------------------------------------------------------------ D. create_synthetic_dataset
    for tag, uprop in (('u10', 0.1), ('u0', 0.0)):
        p2 = configparser.ConfigParser()
        p2.read(os.path.join(args.ref, 'config'))
        p2 = p2['DEFAULT']
        p2['sample_size'] = '23'                                       # 529 voxels: exercises the S^2 % 10 != 0 row drop
        tf.random.set_seed(7)
        tf.random.LOG.clear()
        tx, ty = ref_signals.create_synthetic_dataset(p2, True, True, 0.0, uniform_prop=uprop)
        log = tf.random.LOG
        kinds = [k for k, _ in log]
        assert kinds[:5] == ['uniform', 'normal', 'uniform', 'truncnorm_u', 'perm'], kinds[:5]
        perm = log[4][1].numpy()
        snr_u = np.concatenate([log[5 + 2 * i][1].numpy() for i in range(10)], 0)
        noise = np.concatenate([log[6 + 2 * i][1].numpy() for i in range(10)], 0)
        # marginals exactly as the reference formed them
        ty_np = ty.numpy()
        np.savez(os.path.join(args.out, 'ref_shim_dataset_%s.npz' % tag),
                 train_x=tx.numpy(), train_y=ty_np, perm=perm, snr_u01=snr_u, noise_eps=noise,
                 oef_u01=log[0][1].numpy(), oef_n01=log[1][1].numpy(), dbv_u01=log[2][1].numpy(),
                 dbv_tn_u01=log[3][1].numpy(), uniform_prop=uprop, sample_size=23)
        print('dataset[%s]: x %s y %s' % (tag, tuple(tx.shape), tuple(ty.shape)))

    # ------------------------------------------------------------------ D2. dataset with misalignment / variable Hct
    def small_params():
        p2 = configparser.ConfigParser()
        p2.read(os.path.join(args.ref, 'config'))
        p2 = p2['DEFAULT']
        p2['sample_size'] = '23'
        return p2

    tf.random.set_seed(8)
    tf.random.LOG.clear()
    prob = 0.3
    tx, ty = ref_signals.create_synthetic_dataset(small_params(), True, True, prob, uniform_prop=0.1)
    log = tf.random.LOG
    kinds = [k for k, _ in log]
    assert kinds[:5] == ['uniform', 'normal', 'uniform', 'truncnorm_u', 'perm'], kinds[:5]
    assert kinds[5:11] == ['uniform', 'uniform_int', 'normal', 'normal', 'uniform', 'normal'], kinds[5:11]
    cat = lambda j: np.concatenate([log[5 + 6 * i + j][1].numpy() for i in range(10)], 0)
    np.savez(os.path.join(args.out, 'ref_shim_dataset_misalign.npz'),
             train_x=tx.numpy(), train_y=ty.numpy(), perm=log[4][1].numpy(), prob=prob,
             mis_u01=cat(0).reshape(-1), mis_index=cat(1).reshape(-1).astype(np.int32),
             mis_eps=np.concatenate([cat(2), cat(3)], -1), snr_u01=cat(4), noise_eps=cat(5), sample_size=23)
    print('dataset[misalign]: x %s, %d of %d rows misaligned' % (tuple(tx.shape), int((cat(0) < prob).sum()), tx.shape[0]))

    tf.random.set_seed(9)
    tf.random.LOG.clear()
    tx, ty = ref_signals.create_synthetic_dataset(small_params(), True, True, 0.0, variable_hct=True, uniform_prop=0.1)
    log = tf.random.LOG
    kinds = [k for k, _ in log]
    assert kinds[:6] == ['uniform', 'normal', 'uniform', 'truncnorm_u', 'uniform', 'perm'], kinds[:6]
    np.savez(os.path.join(args.out, 'ref_shim_dataset_hct.npz'),
             train_x=tx.numpy(), train_y=ty.numpy(), perm=log[5][1].numpy(),
             snr_u01=np.concatenate([log[6 + 2 * i][1].numpy() for i in range(10)], 0),
             noise_eps=np.concatenate([log[7 + 2 * i][1].numpy() for i in range(10)], 0), sample_size=23)
    print('dataset[variable_hct]: x %s y %s' % (tuple(tx.shape), tuple(ty.shape)))

    # ------------------------------------------------------------------ A2. variable Hct: forward + tape.gradient
    n3 = 96
    x3 = np.stack([rng.uniform(0.04, 0.84, n3), rng.uniform(0.001, 0.201, n3), rng.uniform(0.2, 0.55, n3)], -1)
    x3 = x3.astype(np.float32)
    x3[:2] = [[0.4, 0.12, 0.34], [0.4, 0.03, 0.45]]
    g3 = rng.standard_normal((n3, 11)).astype(np.float32)
    out3 = {'oef_dbv_hct': x3, 'g_rand': g3}
    for full in (True, False):
        for blood in (True, False):
            layer = ref_signals.SignalGenerationLayer(params, full, blood, variable_hct=True)
            key = 'f%d_b%d' % (int(full), int(blood))
            inp = tf.convert_to_tensor(x3.reshape(n3, 1, 1, 1, 3))
            with tf.GradientTape(persistent=True) as tape:
                tape.watch(inp)
                o3 = layer(inp)
                o3_w = o3 * tf.convert_to_tensor(g3.reshape(n3, 1, 1, 1, 11))
            out3['signal_' + key] = o3.detach().numpy().reshape(n3, 11)
            out3['grad_ones_' + key] = tape.gradient(o3, inp).numpy().reshape(n3, 3)
            out3['grad_rand_' + key] = tape.gradient(o3_w, inp).numpy().reshape(n3, 3)
    # misalignment on the layer itself with a variable-Hct input (signals.py:80-96 with hct [N,1])
    tf.random.set_seed(10)
    tf.random.LOG.clear()
    layer = ref_signals.SignalGenerationLayer(params, True, True, misaligned_prob=0.4, variable_hct=True)
    o3 = layer(tf.convert_to_tensor(x3))
    lg = tf.random.LOG
    assert [k for k, _ in lg] == ['uniform', 'uniform_int', 'normal', 'normal']
    out3.update(mis_prob=0.4, mis_u01=lg[0][1].numpy().reshape(-1), mis_index=lg[1][1].numpy().reshape(-1).astype(np.int32),
                mis_eps=np.concatenate([lg[2][1].numpy(), lg[3][1].numpy()], -1), signal_misaligned=o3.numpy())
    np.savez(os.path.join(args.out, 'ref_shim_forward_hct.npz'), **out3)
    print('forward[variable_hct]: demo input ->', out3['signal_f1_b1'][0])

    # ------------------------------------------------------------------ F. losses either side of the path
    def make_trainer(**kw):
        base = dict(system_params=params, no_units=60, no_intermediate_layers=2, student_t_df=200,
                    initial_im_sigma=0.05, multi_image_normalisation=False, channelwise_gating=True,
                    infer_inv_gamma=False, use_population_prior=False, use_mvg=True, predict_log_data=False)
        base.update(kw)
        return ref_model.EncoderTrainer(**base)

    r = np.random.default_rng(77)
    shp = (2, 6, 5, 3)
    nv = int(np.prod(shp))
    adj = {}
    q5 = np.stack([r.normal(-0.3, 0.9, nv), r.normal(0.0, 0.6, nv), r.normal(-1.2, 0.9, nv),
                   r.normal(0.0, 0.6, nv), r.normal(0.0, 0.8, nv)], -1).astype(np.float32)
    q5[7, 0] = q5[8, 0]                                      # an exact tie: |d| has zero sub-gradient there
    prior5 = (q5 + r.normal(0, 0.3, (nv, 5))).astype(np.float32)
    mask = (r.uniform(size=nv) > 0.3).astype(np.float32)
    labels = np.stack([r.uniform(0.05, 0.8, nv), r.uniform(0.003, 0.195, nv), r.uniform(0.5, 20.0, nv)], -1)
    labels = labels.astype(np.float32)
    labels[0, :2] = [0.04, 0.201]                            # exactly on the transform bounds: exercises the clip
    adj.update(q5=q5, prior5=prior5, mask=mask, labels=labels, shape=np.array(shp))

    def t5(a, c):
        t = tf.convert_to_tensor(a.reshape(shp + (c,)))
        t.requires_grad_(True)
        return t

    m_t = tf.convert_to_tensor(mask.reshape(shp + (1,)))
    # F1 smoothness (model.py:726-754), mvg and diagonal layouts
    tr = make_trainer()
    q_t = t5(q5, 5)
    tv = tr.smoothness_loss(tf.concat([tf.convert_to_tensor(prior5.reshape(shp + (5,))), m_t], -1), q_t)
    adj['tv_mvg'] = tv.detach().numpy()
    adj['tv_mvg_grad'] = torch.autograd.grad(tv, [q_t])[0].numpy().reshape(nv, 5)
    tr_d = make_trainer(use_mvg=False)
    q4_t = t5(q5[:, :4], 4)
    tv = tr_d.smoothness_loss(tf.concat([tf.convert_to_tensor(prior5[:, :4].reshape(shp + (4,))), m_t], -1), q4_t)
    adj['tv_diag'] = tv.detach().numpy()
    adj['tv_diag_grad'] = torch.autograd.grad(tv, [q4_t])[0].numpy().reshape(nv, 4)
    # F2 pre-training loss (model.py:449-514): mvg / diagonal, with and without the inverse-gamma term
    lab_t = tf.convert_to_tensor(labels.reshape(shp + (3,)))
    for tag, trn, c, ig in (('mvg', tr, 5, (0.0, 0.0)), ('mvg_ig', tr, 5, (3.0, 0.15)),
                            ('diag', tr_d, 4, (0.0, 0.0)), ('diag_ig', tr_d, 4, (3.0, 0.15))):
        lab = labels.copy()
        if c == 4:
            lab[0, :2] = [0.3, 0.05]                         # the diagonal branch has no clip: keep labels interior
        # With the inverse-gamma term the reference subtracts a [N] vector from the [B,X,Y,Z] loss (model.py:507),
        # which only broadcasts for [N,1,1,1] inputs (to [N,1,1,N]; its mean = mean(loss) - mean(prior term)).
        vshape = (nv, 1, 1, 1) if ig[0] * ig[1] > 0.0 else shp
        q_t = tf.convert_to_tensor(q5[:, :c].reshape(vshape + (c,)))
        q_t.requires_grad_(True)
        loss = trn.synthetic_data_loss(tf.convert_to_tensor(lab.reshape(vshape + (3,))), q_t, False, ig[0], ig[1])
        adj['synth_%s' % tag] = loss.detach().numpy()
        adj['synth_%s_grad' % tag] = torch.autograd.grad(loss, [q_t])[0].numpy().reshape(nv, c)
        adj['synth_%s_labels' % tag] = lab
    # F2a' infer_inv_gamma (model.py:454-455,493-496): learned InverseGamma parameters ride in 4 extra channels
    tr_i = make_trainer(use_mvg=False, infer_inv_gamma=True)
    lab = labels.copy()
    lab[0, :2] = [0.3, 0.05]
    ig = np.array([18.0, 2.2, 23.0, 2.9], np.float32)
    q8 = np.concatenate([q5[:, :4], np.tile(ig[None, :], (nv, 1))], -1).astype(np.float32)
    q_t = tf.convert_to_tensor(q8.reshape(nv, 1, 1, 1, 8))
    q_t.requires_grad_(True)
    loss = tr_i.synthetic_data_loss(tf.convert_to_tensor(lab.reshape(nv, 1, 1, 1, 3)), q_t, False, 0.0, 0.0)
    adj['synth_iginf'] = loss.detach().numpy()
    adj['synth_iginf_grad'] = torch.autograd.grad(loss, [q_t])[0].numpy().reshape(nv, 8)
    adj['synth_iginf_pred'] = q8
    adj['synth_iginf_labels'] = lab
    # F2b r2p term (10 recorded reparam draws, model.py:480-494)
    tf.random.set_seed(21)
    tf.random.LOG.clear()
    q_t = tf.convert_to_tensor(q5.reshape(nv, 1, 1, 1, 5))           # same broadcast constraint (model.py:490)
    q_t.requires_grad_(True)
    loss = tr.synthetic_data_loss(tf.convert_to_tensor(labels.reshape(nv, 1, 1, 1, 3)), q_t, True, 0.0, 0.0)
    adj['synth_r2p'] = loss.detach().numpy()
    adj['synth_r2p_grad'] = torch.autograd.grad(loss, [q_t])[0].numpy().reshape(nv, 5)
    adj['synth_r2p_eps'] = np.stack([e[1].numpy().reshape(nv, 2) for e in tf.random.LOG], 1)
    # F3 metrics (model.py:345-374): 20 recorded draws each
    for i, name in enumerate(('oef', 'dbv', 'r2p')):
        tf.random.set_seed(30 + i)
        tf.random.LOG.clear()
        val = getattr(tr, name + '_metric')(lab_t, tf.convert_to_tensor(q5.reshape(shp + (5,))))
        adj['metric_' + name] = val.detach().numpy()
        adj['metric_%s_eps' % name] = np.stack([e[1].numpy().reshape(nv, 2) for e in tf.random.LOG], 1)
    # F4 diagonal KL variants (model.py:685-716): plain and population prior
    q4_t = t5(q5[:, :4], 4)
    kl = tr_d.kl_loss(tf.concat([tf.convert_to_tensor(prior5[:, :4].reshape(shp + (4,))), m_t], -1), q4_t)
    adj['kl_diag'] = kl.detach().numpy()
    adj['kl_diag_grad'] = torch.autograd.grad(kl, [q4_t])[0].numpy().reshape(nv, 4)
    tr_p = make_trainer(use_mvg=False, use_population_prior=True)
    pop = np.tile(np.array([[-0.2, 0.3, -1.0, 0.2]], np.float32), (nv, 1))
    q8_t = t5(np.concatenate([q5[:, :4], pop], -1), 8)
    kl = tr_p.kl_loss(tf.concat([tf.convert_to_tensor(prior5[:, :4].reshape(shp + (4,))), m_t], -1), q8_t)
    adj['kl_pop'] = kl.detach().numpy()
    adj['kl_pop_grad'] = torch.autograd.grad(kl, [q8_t])[0].numpy().reshape(nv, 8)
    adj['kl_pop_pred'] = q8_t.detach().numpy().reshape(nv, 8)
    # F5 mixture-of-Gaussians population prior (model.py:666-684): q + 3 components, two recorded normal draws
    tr_m = make_trainer(use_mvg=False, use_population_prior=True, mog_components=3)
    comps = np.tile(np.array([[-0.4, 0.2, -1.3, 0.1, 0.3, -0.2, -0.8, 0.3, -1.0, 0.4, -1.6, -0.1]], np.float32), (nv, 1))
    q16_t = t5(np.concatenate([q5[:, :4], comps], -1), 16)
    tf.random.set_seed(41)
    tf.random.LOG.clear()
    kl = tr_m.kl_loss(tf.concat([tf.convert_to_tensor(prior5[:, :4].reshape(shp + (4,))), m_t], -1), q16_t)
    adj['kl_mog'] = kl.detach().numpy()
    adj['kl_mog_grad'] = torch.autograd.grad(kl, [q16_t])[0].numpy().reshape(nv, 16)
    adj['kl_mog_pred'] = q16_t.detach().numpy().reshape(nv, 16)
    adj['kl_mog_eps'] = np.stack([e[1].numpy().reshape(nv) for e in tf.random.LOG], -1)      # [nv, 2]: oef, dbv draws
    np.savez(os.path.join(args.out, 'ref_shim_adjacent.npz'), **adj)
    print('adjacent: tv=%.6f synth=%.6f kl_diag=%.6f kl_pop=%.6f' % (float(adj['tv_mvg']), float(adj['synth_mvg']),
                                                                   float(adj['kl_diag']), float(adj['kl_pop'])))

    # ------------------------------------------------------------------ G. the amortization network itself
    # EncoderTrainer.create_encoder (model.py:122-223) executed over the shim's minimal Keras functional API: outputs
    # of the outer model and the gradient of a weighted sum of them w.r.t. every kernel / bias.
    tr_e = ref_model.EncoderTrainer(system_params=params, no_units=60, no_intermediate_layers=2, student_t_df=200,
                                    initial_im_sigma=0.05, activation_type='relu', multi_image_normalisation=False,
                                    channelwise_gating=True, infer_inv_gamma=False, use_population_prior=False,
                                    use_mvg=True, predict_log_data=False)
    torch.manual_seed(2026)
    del tf.keras.layers.created[:]
    outer, _inner = tr_e.create_encoder(gate_offset=-1.0, resid_init_std=0.1, no_ip_images=11)
    r = np.random.default_rng(99)
    vol = (2, 6, 5, 3)
    e_data = (r.uniform(20.0, 70.0, vol + (11,))).astype(np.float32)
    outer(tf.convert_to_tensor(e_data))                                   # builds the weights (creation order below)
    layers = list(tf.keras.layers.created)
    assert len(layers) == 11, len(layers)            # first | (pointwise, conv_a, conv_b, gate) x 2 | final | im_sigma
    for lay in layers:
        lay.bias = (lay.bias + 0.05 * torch.randn_like(lay.bias))         # non-trivial biases
        lay.kernel.requires_grad_(True)
        lay.bias.requires_grad_(True)
    o0, o1, o2 = outer(tf.convert_to_tensor(e_data))
    e_w = [r.standard_normal(tuple(o.shape)).astype(np.float32) for o in (o0, o1, o2)]
    total = sum((o * torch.as_tensor(w)).sum() for o, w in zip((o0, o1, o2), e_w))
    grads = torch.autograd.grad(total, [t for lay in layers for t in (lay.kernel, lay.bias)])
    enc_fix = {'data': e_data, 'out_voxelwise': o0.detach().numpy(), 'out_spatial': o1.detach().numpy(),
               'out_sigma': o2.detach().numpy(), 'gate_offset': -1.0, 'n_layers': len(layers)}
    for i, w in enumerate(e_w):
        enc_fix['w_out%d' % i] = w
    for i, lay in enumerate(layers):
        enc_fix['kernel%d' % i] = lay.kernel.detach().numpy()              # keras layout [kx, ky, kz, C_in, C_out]
        enc_fix['bias%d' % i] = lay.bias.detach().numpy()
        enc_fix['grad_kernel%d' % i] = grads[2 * i].numpy()
        enc_fix['grad_bias%d' % i] = grads[2 * i + 1].numpy()
    np.savez(os.path.join(args.out, 'ref_shim_encoder.npz'), **enc_fix)
    print('encoder: %d layers, %d parameters, sigma mean %.4f' % (
        len(layers), sum(int(np.prod(lay.kernel.shape)) + int(lay.bias.numel()) for lay in layers),
        float(o2.detach().mean())))

    # ------------------------------------------------------------------ E. Appendix B KATs
    np.savez(os.path.join(args.out, 'kat_appendix_b.npz'),
             oef_dbv=np.array([[0.4, 0.12], [0.4, 0.03]]),
             signal_fp64_0=np.array([0.37586412, 0.40909051, 0.42242562, 0.40909051, 0.37586412, 0.33616669,
                                     0.29923692, 0.26730955, 0.23914064, 0.21358386, 0.19047615]),
             grad_sum_0=np.array([-3.06411134, -7.29813321]),
             signal_fp32_1=np.array([0.41349795, 0.42242014, 0.4258472, 0.42242014, 0.41349795, 0.4020199,
                                     0.39038378, 0.3794184, 0.36889765, 0.35852298, 0.34832814]),
             tissue_tau0=np.array(0.42698773))


if __name__ == '__main__':
    main()

"""Step the CUDA engine one iteration at a time and compare every intermediate buffer with the oracle trace.
Debug aid (run on the GPU box):  python tools/debug_rollout.py [case] [max_iters]"""
import os, sys
os.environ['INFGEN_NO_EARLY_EDGES'] = '1'     # keep the edge buffers of iteration t readable after iteration t
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.golden.cases import build_case
from oracle.agent_decoder_oracle import rollout
from infgen_b200.agent_decoder import B200AgentDecoder
from infgen_b200.host import prepare_scene, HostBatch


def rep(name, got, want, tol=1e-4):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    if got.shape != want.shape:
        print(f'  {name}: SHAPE {got.shape} vs {want.shape}')
        return
    err = np.abs(got - want)
    bad = err > tol + 1e-3 * np.abs(want)
    flag = 'ok ' if not bad.any() else 'BAD'
    print(f'  [{flag}] {name}: max err {err.max() if err.size else 0:.3e} bad {bad.sum()}/{bad.size}'
          + (f' first {np.argwhere(bad)[0].tolist()} got {got[bad][0]:.6g} want {want[bad][0]:.6g}' if bad.any() else ''))


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else 'cfg0_a8'
    max_iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    scene, sd, cfg, spec = build_case(case)
    o = rollout(scene, sd, cfg, collect_trace=True, assume_no_insertion=True, max_iters=max_iters)
    tr = o['trace']
    dec = B200AgentDecoder(sd, cfg, use_cuda_graph=False, trace=True)
    sh = prepare_scene(scene, scene['map_enc'], cfg)
    b = HostBatch([sh], cfg, [0])
    dec.load(b, [sh])
    A, R, T, HC, cap = sh.n_rows, b.R, b.T, cfg.hist_cols, b.cap
    # force the oracle's tokens / states so divergence cannot compound
    S = sh.n_iters
    ftok = torch.zeros(R, S, dtype=torch.int32); fst = torch.ones(R, S, dtype=torch.int32)
    for t, it in enumerate(tr):
        ftok[:A, t] = it['token'].int(); fst[:A, t] = it['state'].int()
    dec.set_forcing(ftok, fst)
    dec.prefill()
    for t, it in enumerate(tr):
        print(f'iteration {t} (cur={it["cur"]})')
        dec.step(1)
        dec.synchronize()
        cur, nxt = it['cur'], it['cur'] + 1
        # edges (the buffers still hold this iteration's lists)
        for key, cnt_name, stride in (('edges_t', 't_cnt', cfg.window), ('edges_m', 'm_cnt', cfg.max_pl2a_neighbors)):
            e = it[key]
            cnt = dec.debug_read(cnt_name, (R,), np.int32)[:A]
            div = T if key == 'edges_t' else A
            dst_local = (e['dst'] // T) if key == 'edges_t' else (e['dst'] % A)
            want_cnt = np.bincount(dst_local.numpy(), minlength=A)
            rep(f'{cnt_name}', cnt, want_cnt, 0)
            D = e['raw'].shape[1]
            raw = dec.debug_read(cnt_name[0] + '_raw', (R * stride, 4 if key == 'edges_t' else 3))
            rhat = dec.debug_read('rhat_' + cnt_name[0], (R * stride, 128))
            if np.array_equal(cnt, want_cnt) and e['raw'].shape[0]:
                sel = np.concatenate([np.arange(a * stride, a * stride + cnt[a]) for a in range(A)])
                order = torch.argsort(dst_local, stable=True).numpy()
                rep(f'{key} raw', raw[sel][:, :D], e['raw'].numpy()[order])
                emb = e['emb'][order]
                embn = (emb - emb.mean(1, keepdim=True)) / torch.sqrt(emb.var(1, unbiased=False, keepdim=True) + 1e-5)
                rep(f'{key} rhat', rhat[sel], embn.numpy(), 1e-3)
        e = it['edges_a']
        cnt = dec.debug_read('a_cnt', (R,), np.int32)[:A]
        start = dec.debug_read('a_start', (R,), np.int32)[:A]
        dst_local = e['dst'] % A
        want_cnt = np.bincount(dst_local.numpy(), minlength=A)
        rep('a_cnt', cnt, want_cnt, 0)
        if np.array_equal(cnt, want_cnt) and e['raw'].shape[0]:
            raw = dec.debug_read('a_raw', (R * cap, 3)); rhat = dec.debug_read('rhat_a', (R * cap, 128))
            src = dec.debug_read('a_src', (R * cap,), np.int32)
            sel = np.concatenate([np.arange(start[a], start[a] + cnt[a]) for a in range(A)])
            order = torch.argsort(dst_local, stable=True).numpy()
            rep('edges_a src', src[sel], (e['src'] % A).numpy()[order], 0)
            rep('edges_a raw', raw[sel], e['raw'].numpy()[order])
            emb = e['emb'][order]
            embn = (emb - emb.mean(1, keepdim=True)) / torch.sqrt(emb.var(1, unbiased=False, keepdim=True) + 1e-5)
            rep('edges_a rhat', rhat[sel], embn.numpy(), 1e-3)
        ta = dec.trace_arrays()
        for i in range(6):
            rep(f'layer_out[{i}]', ta['layer_out'][t, i, :A], it['layer_out'][i].numpy(), 1e-3)
        rep('head_in', ta['head_in'][t, :A], it['head_in'].numpy(), 1e-3)
        rep('token_logits', ta['token_logits'][t, :A], it['token_logits'].numpy(), 1e-3)
        rep('state_logits', ta['state_logits'][t, :A], it['state_logits'].numpy(), 1e-3)
        pos = dec.debug_read('pos', (R, T, 2)); head = dec.debug_read('head', (R, T))
        grid = dec.debug_read('grid', (R, T), np.int32); state = dec.debug_read('state', (R, T), np.int32)
        token = dec.debug_read('token', (R, T), np.int32)
        rep('pos_next', pos[:A, nxt], it['pos_next'].numpy())
        rep('head_next', head[:A, nxt], it['head_next'].numpy())
        rep('grid_next', grid[:A, nxt], it['grid_next'].numpy(), 0)
        rep('state_next', state[:A, nxt], it['state'].numpy(), 0)
        rep('token_next', token[:A, nxt], it['token'].numpy(), 0)
        rep('feat_next (x)', dec.debug_read('x', (R, 128))[:A], it['feat_next'].numpy(), 1e-3)
        # pieces of the column embedding, recomputed with the oracle's functions from the engine's own state
        from oracle.agent_decoder_oracle import (mlp_embedding, fourier_embedding, build_grid, angle_between)
        import math
        P, Hd, St = torch.from_numpy(pos[:A]), torch.from_numpy(head[:A]), torch.from_numpy(state[:A]).long()
        mv = P[:, nxt] - P[:, cur]
        hv = torch.stack([Hd[:, nxt].cos(), Hd[:, nxt].sin()], -1)
        feat = torch.stack([mv.norm(p=2, dim=-1), angle_between(hv, mv)], -1)
        rep('xa_raw', dec.debug_read('xa_raw', (R, 2))[:A], feat.numpy())
        ag = scene['agent']
        filt = ag['state_idx'][:, HC - 1] != 0
        type_a = ag['type'][filt].long(); shp = ag['shape'][filt][:, cfg.num_historical_steps - 1]
        cat = sd['type_a_emb.weight'][type_a] + mlp_embedding(sd, 'shape_emb', shp)
        rep('cat_tab', dec.debug_read('cat_tab', (R + 1, 128))[:A], cat.numpy())
        # generated columns carry the seed-type / 0.1-shape categorical row
        rep('cat_idx', dec.debug_read('cat_idx', (R,), np.int32)[:A], np.full(A, R), 0)
        cat_seed = sd['type_a_emb.weight'][3][None] + mlp_embedding(sd, 'shape_emb', torch.full((1, 3), 0.1))
        xa = fourier_embedding(sd, 'x_a_emb', feat, cat_seed.expand(A, 128))
        # (x_a_emb itself stays on chip inside k_embed_column; it is checked through the fused feature below)
        vocab = torch.stack([ag['trajectory_token_veh'], ag['trajectory_token_ped'], ag['trajectory_token_cyc']])
        tok_row = dec.debug_read('tok_row', (R,), np.int32)[:A]
        rep('tok_row', tok_row, (type_a * 2050 + torch.from_numpy(token[:A, nxt]).long()).numpy(), 0)
        rep('state_idx', dec.debug_read('state_idx', (R,), np.int32)[:A], state[:A, nxt], 0)
        rep('grid_row', dec.debug_read('grid_row', (R,), np.int32)[:A], grid[:A, nxt], 0)
        tok_e = torch.stack([mlp_embedding(sd, f'token_emb_{nm}', vocab[ti][:, -1].flatten(1, 2))[token[a, nxt]]
                             for a, (ti, nm) in enumerate([(int(x), ('veh', 'ped', 'cyc')[int(x)]) for x in type_a])])
        g = build_grid()
        grid_e = mlp_embedding(sd, 'token_emb_grid', g)[torch.from_numpy(grid[:A, nxt]).long()]
        st_e = sd['state_a_emb.weight'][St[:, nxt]]
        fus = mlp_embedding(sd, 'fusion_emb', torch.cat([tok_e, xa, st_e, grid_e], -1))
        rep('fusion recomputed vs oracle feat_next', fus.numpy(), it['feat_next'].numpy(), 1e-3)
        rep('x vs fusion recomputed', dec.debug_read('x', (R, 128))[:A], fus.numpy(), 1e-3)
    dec.close()


if __name__ == '__main__':
    main()

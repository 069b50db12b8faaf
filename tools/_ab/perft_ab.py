"""A/B of mc_perft between two builds of libmcaz.so on one box (scratch)."""
import ctypes, os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
STATE = np.dtype([('pl0', '<u4'), ('pl1', '<u4'), ('pl2', '<u4'), ('white', '<u4'), ('meta', '<u4')])
for name, path in (('old', os.path.join(REPO, 'tools/_ab/libmcaz_old.so')), ('new', os.path.join(REPO, 'minitchess_alphazero_b200/libmcaz.so')),
                   ('old', os.path.join(REPO, 'tools/_ab/libmcaz_old.so')), ('new', os.path.join(REPO, 'minitchess_alphazero_b200/libmcaz.so'))):
    L = ctypes.CDLL(path)
    L.mcaz_set_device(0)
    st = np.zeros(1, STATE)
    L.mc_state_from_fen(b'2nbk/2ppp/5/5/PPP2/KBN2 w 0 1', st.ctypes.data_as(ctypes.c_void_p))
    nodes = np.zeros(1, np.uint64)
    line = []
    for depth in (5, 6, 7, 7, 8):
        t0 = time.perf_counter()
        rc = L.mc_perft(st.ctypes.data_as(ctypes.c_void_p), 1, depth, None, nodes.ctypes.data_as(ctypes.c_void_p))
        line.append('d%d %d nodes %.2f ms rc %d' % (depth, int(nodes[0]), (time.perf_counter() - t0) * 1e3, rc))
    print(name, ' | '.join(line), flush=True)

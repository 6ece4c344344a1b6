import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import __graft_entry__ as ge
from oracle import synth, restate as R
from test_gpu_parity import _expected_totals
pkg = ge.load_package()
np.set_printoptions(linewidth=200, precision=6)
for motion in (0, 1, 2, 3):
    w, h = 320, 240
    st_ = synth.Stack(w, h, 2, motion, seed=20 + motion)
    f0, f1 = st_.frames()
    rng = np.random.default_rng(motion)
    g = st_.truth[1].copy(); g[:2, 2] += rng.uniform(-1.5, 1.5, 2)
    m32 = g.astype(np.float32)
    if motion != 3: m32[2] = (0, 0, 1)
    if motion == 1:
        th = math.asin(float(m32[1, 0])); m32[0, 0] = m32[1, 1] = np.float32(math.cos(th)); m32[0, 1] = -m32[1, 0]
    params = pkg.EccMatchParameters(pkg.MotionType(motion), 50, 1e-5, 5)
    with pkg.EccStack(w, h, 3, params, device=0, lanes=1) as st:
        st.set_reference(f0)
        tot, m_out, rho, status = st.debug_iteration(f1, m32)
    tmpl = R.gaussian_blur_f32(R.bgr2gray_u8(f1).astype(np.float32), 5)
    img = R.gaussian_blur_f32(R.bgr2gray_u8(f0).astype(np.float32), 5)
    mm = m32 if motion == 3 else m32[:2]
    want = _expected_totals(motion, tmpl, img, mm)
    rel = np.abs(tot - want) / np.maximum(np.abs(want), 1e-30)
    print('motion', motion, 'status', status, 'rho', rho)
    print(' m_in', m32.ravel()); print(' m_out', m_out.ravel())
    sums = R.ecc_sums(motion, tmpl, img, *R.central_gradients(img), mm)
    rho_want, m_want = R.ecc_epilogue(motion, sums, mm)
    print(' m_want', np.asarray(m_want).ravel(), 'rho_want', rho_want)
    for i in range(len(tot)):
        flag = '' if rel[i] < 2e-4 else '  <<<<'
        print(f'  {i:3d} got {tot[i]: .9e} want {want[i]: .9e} rel {rel[i]:.2e}{flag}')

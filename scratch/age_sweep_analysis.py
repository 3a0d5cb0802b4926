import sys; sys.path.insert(0,'.')
import numpy as np, scipy.fft as sf, time
from oracle import scarplet_oracle as O
z = np.load('tests/golden/synthetic_dem_f32.npy').astype(np.float64)
gold = np.load('tests/golden/reference_goldens.npz')['synthetic_match1']
ny,nx = z.shape
angles = O.search_angles(); ages = O.default_ages()
def mt(z,kind,scale,age,angle,f32):
    curv = O.directional_laplacian(z,1.,1.,angle)
    t = O.template_array(kind,scale,age,angle,nx,ny,1.)
    M = t!=0; n = M.sum()+O.EPS; ts = (t**2).sum()
    if f32:
        fc = sf.fft2(curv.astype(np.float32)); fc2 = sf.fft2((curv**2).astype(np.float32)); ft = sf.fft2(t.astype(np.float32)); fm = sf.fft2(M.astype(np.float32))
    else:
        fc = sf.fft2(curv); fc2 = sf.fft2(curv**2); ft = sf.fft2(t); fm = sf.fft2(M.astype(float))
    xc = np.real(np.fft.fftshift(sf.ifft2(ft*fc))).astype(np.float64)
    T3 = np.real(np.fft.fftshift(sf.ifft2(fc2*fm))).astype(np.float64)
    amp = xc/ts; T1 = ts*amp**2
    err = (T1-2*amp*xc+T3)/n + O.EPS
    snr = np.abs(T1/err)
    c,d = O.template_constants(kind,scale,age,nx)
    wl = O.window_limits(kind,nx,ny,1.,-angle,c,d)
    amp[wl]=0; snr[wl]=0
    return amp,snr
t0=time.time()
S64 = np.zeros((len(ages),len(angles),ny,nx),np.float32); S32 = np.zeros_like(S64)
for gi,age in enumerate(ages):
    for ai,a in enumerate(angles):
        S64[gi,ai] = mt(z,O.SCARP,100,age,a,False)[1]
        S32[gi,ai] = mt(z,O.SCARP,100,age,a,True)[1]
print('%.0fs'%(time.time()-t0))
np.save('/tmp/S64.npy',S64); np.save('/tmp/S32.npy',S32)

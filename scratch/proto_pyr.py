"""Prototype of the flattened (per-level independent) formulation used by the CUDA kernels."""
import sys, math, numpy as np, torch
sys.path.insert(0,'/root/repo')
from oracle.steerable_shim import SCFpyr_PyTorch, level_sizes, prepare_grid, rcosFn, pointOp, crop_start

def signed(n):  # signed frequency of unshifted index
    k = np.arange(n); return np.where(k < (n+1)//2, k, k-n)

def plan(H,W,height,nb,s):
    sizes = level_sizes(H,W,height,s)
    L = height-2
    log_rad, angle = prepare_grid(H,W)
    Xr, Yr = rcosFn(1,-0.5); Yr = np.sqrt(Yr); YIr = np.sqrt(np.abs(1-Yr**2))
    # radial masks on full shifted grid evaluated lazily per level window
    tabs = []
    r0, c0 = 0, 0
    lr = log_rad
    lo_prod = pointOp(lr, YIr, Xr)  # lo0
    hi0 = pointOp(log_rad, Yr, Xr)
    X = Xr.copy()
    for l in range(L):
        X = X - np.log2(s)
        himask = pointOp(lr, Yr, X)
        D = lo_prod*himask            # shifted layout, level window
        tabs.append(np.fft.ifftshift(D))   # unshifted layout
        h,w = lr.shape
        nh, nw = sizes[l+1]
        sr, sc = crop_start(h,nh), crop_start(w,nw)
        lr = lr[sr:sr+nh, sc:sc+nw]
        lo_prod = lo_prod[sr:sr+nh, sc:sc+nw]*pointOp(lr, YIr, X)
    Dlow = np.fft.ifftshift(lo_prod)
    return sizes, tabs, Dlow, np.fft.ifftshift(hi0)

def angular(h,w,H,W,nb,two_sided):
    fy = signed(h)[:,None]*2.0/H; fx = signed(w)[None,:]*2.0/W
    rad = np.sqrt(fx*fx+fy*fy); rad[0,0]=1.0
    order = nb-1
    const = 2**(2*order)*math.factorial(order)**2/(nb*math.factorial(2*order))
    out=[]
    for b in range(nb):
        phi = math.pi*b/nb
        c = (fx*math.cos(phi)+fy*math.sin(phi))/rad
        c[0,0] = math.cos(phi)
        if two_sided: A = math.sqrt(const)*c**order
        else: A = 2*math.sqrt(const)*c**order*(c>0)
        out.append(A)
    return out

def build(x, height, nb, s):
    N,H,W = x.shape
    sizes,tabs,Dlow,hi0 = plan(H,W,height,nb,s)
    X = np.fft.fft2(x)
    L=height-2
    fac = (-1j)**(nb-1)
    bands=[]
    for l in range(L):
        h,w = sizes[l]
        ky = signed(h)%H; kx = signed(w)%W
        Xl = X[:, ky[:,None], kx[None,:]]*tabs[l]
        A = angular(h,w,H,W,nb,False)
        bands.append([np.fft.ifft2(Xl*A[b]*fac) for b in range(nb)])
    h,w = sizes[L]
    ky = signed(h)%H; kx = signed(w)%W
    low = np.fft.ifft2(X[:, ky[:,None], kx[None,:]]*Dlow).real
    high = np.fft.ifft2(X*hi0).real
    return high,bands,low

def reconstruct(high,bands,low,height,nb,s):
    N,H,W = high.shape
    sizes,tabs,Dlow,hi0 = plan(H,W,height,nb,s)
    L=height-2
    S = np.fft.fft2(high)*hi0
    fac = (1j)**(nb-1)
    for l in range(L):
        h,w = sizes[l]
        A = angular(h,w,H,W,nb,True)
        Y = sum(np.fft.fft2(bands[l][b])*A[b] for b in range(nb))*fac*tabs[l]
        ky = signed(h)%H; kx = signed(w)%W
        S[:, ky[:,None], kx[None,:]] += Y
    h,w = sizes[L]
    ky = signed(h)%H; kx = signed(w)%W
    S[:, ky[:,None], kx[None,:]] += np.fft.fft2(low)*Dlow
    return np.fft.ifft2(S).real

if __name__=="__main__":
    for (H,W,height) in [(256,256,12),(90,150,8),(135,241,9)]:
        s=np.sqrt(2); nb=4
        x = torch.rand(2,1,H,W, dtype=torch.float64)
        pyr = SCFpyr_PyTorch(height=height,nbands=nb,scale_factor=s,precision="fp64")
        pyr.rdtype=torch.float64
        c = pyr.build(x)
        high,bands,low = build(x.squeeze(1).numpy(),height,nb,s)
        print("high",np.abs(high-c[0].numpy()).max(),"low",np.abs(low-c[-1].numpy()).max()/np.abs(low).max())
        for l in range(height-2):
            e = max(np.abs(bands[l][b]-torch.view_as_complex(c[1+l][b]).numpy()).max() for b in range(nb))
            m = max(np.abs(bands[l][b]).max() for b in range(nb))
            print(" level",l,"abs err",e,"rel",e/m)
        # reconstruct from shim coeffs (perturbed so spectrum isn't one-sided)
        rng=np.random.default_rng(0)
        c2=[c[0]]+[[bb*torch.from_numpy(rng.uniform(0.5,1.5,bb.shape)).float() for bb in lv] for lv in c[1:-1]]+[c[-1]]
        r_ref = pyr.reconstruct(c2).numpy()
        b2 = [[torch.view_as_complex(bb.contiguous()).numpy().astype(np.complex128) for bb in lv] for lv in c2[1:-1]]
        r = reconstruct(c2[0].numpy().astype(np.float64), b2, c2[-1].numpy().astype(np.float64), height,nb,s)
        print(" recon err", np.abs(r-r_ref).max())

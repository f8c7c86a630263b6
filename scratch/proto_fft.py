import numpy as np
def factorize(n):
    f=[]; 
    # large primes first (generic), then 5,3,4,2 ... order: we put generic/odd first, radix-4 last
    m=n; small=[]
    for p in [4,2,3,5]:
        pass
    fac=[]
    d=2; rem=n; primes=[]
    while d*d<=rem:
        while rem%d==0: primes.append(d); rem//=d
        d+=1
    if rem>1: primes.append(rem)
    twos = primes.count(2); others=sorted([p for p in primes if p!=2], reverse=True)
    fac = others + [4]*(twos//2) + [2]*(twos%2)
    return fac
def stockham(x, inv=False):
    n=len(x); fac=factorize(n); tw=np.exp((2j if inv else -2j)*np.pi*np.arange(n)/n)
    a=x.astype(np.complex128).copy(); Ns=1
    for r in fac:
        m=n//r; step=n//(Ns*r); b=np.zeros_like(a)
        for o in range(n):                 # output-centric generic stage
            k=o%Ns; v=(o//Ns)%r; jhi=o//(Ns*r); j=jhi*Ns+k
            acc=0; e=0; t=0
            for u in range(r):
                idx=(e+t*m)%n
                acc+=a[j+u*m]*tw[idx]
                e+=k*step; t+=v
                if t>=r: t-=r
            b[o]=acc
        a=b; Ns*=r
    return a
for n in [8,12,45,91,181,340,764,1358//2]:
    x=np.random.rand(n)+1j*np.random.rand(n)
    print(n, factorize(n), np.abs(stockham(x)-np.fft.fft(x)).max(), np.abs(stockham(x,True)-np.fft.ifft(x)*n).max())

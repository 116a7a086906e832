"""Developer script (GPU): L2 properties of device 0 (persisting carve-out, access-policy window limit)."""
from cuda.bindings import runtime as rt
err, p = rt.cudaGetDeviceProperties(0)
print("l2CacheSize %.1f MB, persistingL2CacheMaxSize %.1f MB, accessPolicyMaxWindowSize %.1f MB" % (p.l2CacheSize / 2**20, p.persistingL2CacheMaxSize / 2**20, p.accessPolicyMaxWindowSize / 2**20))

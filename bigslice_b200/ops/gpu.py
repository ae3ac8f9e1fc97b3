"""placeholder until the ctypes bindings are written"""
HASHABLE = ()
SEGREDUCE_DTYPES = ()
def available():
    return False

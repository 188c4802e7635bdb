"""Import shim so the unmodified reference `model` package imports without the external
`fcos_core` extension (tianzhi0549/FCOS, unpinned). Test infrastructure only."""


class _Missing:
    def __getattr__(self, name):
        def _raise(*a, **k):
            raise RuntimeError("fcos_core._C.%s is not available in the oracle shim" % name)
        return _raise


_C = _Missing()

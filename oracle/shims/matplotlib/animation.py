"""TEST INFRASTRUCTURE ONLY -- import-only matplotlib submodule stub (render() is out of scope)."""


class _Stub:
    def __init__(self, *a, **k):
        raise NotImplementedError("matplotlib is stubbed in the oracle shims; rendering is out of scope")


MarkerStyle = FormatStrFormatter = PillowWriter = Polygon = FuncAnimation = _Stub

"""Minimal stand-in for `gymnasium` (absent from this image) so that the reference
package can be imported by tests/golden/make_golden.py. Test infrastructure only."""


class Env:
    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)


class ObservationWrapper(Wrapper):
    pass


from . import spaces, wrappers  # noqa: E402,F401

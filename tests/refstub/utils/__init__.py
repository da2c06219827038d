from .ops import *  # noqa: F401,F403  (the reference's utils/__init__.py re-exports its ops the same way)

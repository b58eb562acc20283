"""veros_b200: sm_100a CUDA implementation of Veros's isoneutral mixing step and column solve,
behind the reference's own call surface (see DESIGN.md / INTEGRATION.md)."""

__all__ = ["isoneutral", "utilities", "state"]

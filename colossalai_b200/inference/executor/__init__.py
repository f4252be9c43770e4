from .rpc_worker import RPCWorker, serve_worker

__all__ = ["RPCWorker", "serve_worker"]

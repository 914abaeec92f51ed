"""Training runtime for AdvancedNCF on B200.

`NCFTrainEngine` is the fast path: one call = one iteration of the reference's
`ModelTrainer.train_epoch` loop body (trainer.py:253-289: forward, BCELoss, zero_grad, backward,
Adam) executed entirely by libncf_b200.so (`ncf_train_step`), with the dense Adam state in one
flat buffer and the table update fused into the backward.

`ModelTrainer` mirrors the reference class (same constructor arguments, `train_epoch`, `validate`,
`train`) for callers that drive the model through `model(features)` / `loss.backward()` /
`optimizer.step()`; BigQuery loading (trainer.py:164-205) is out of scope - loaders are passed in.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib
from .architecture import AdvancedNCF, _stream
from .metrics import calculate_metrics


class NCFTrainEngine:
    def __init__(self, model: AdvancedNCF, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-5, table_mode: str = "fused_dense_equiv", max_rows: int = 0):
        if table_mode not in ("fused_dense_equiv", "fused_sparse"):
            raise ValueError("the engine updates the tables itself: table_mode must be fused_dense_equiv or fused_sparse")
        self.model = model
        self.lib = _lib.load()
        model._ensure_flat()
        model.configure_table_optimizer(table_mode, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        dev = model._flat.device
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        n = model._flat.numel()
        self.dense_grad = torch.zeros(n, device=dev)
        self.dense_m = torch.zeros(n, device=dev)
        self.dense_v = torch.zeros(n, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.hp = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.table_mode = table_mode
        self.step = 0
        self._ws = None
        self._out = None
        self._dev_in = None
        self._model_sig = None
        self._dense_ends = None
        self._tables = None
        self.S = 1 + model.negative_samples
        # the id sort of the embedding backward runs on this stream, next to the forward (ncf_set_aux_stream)
        self._aux_stream = torch.cuda.Stream(device=dev) if os.environ.get("NCF_AUX_STREAM", "1") != "0" else None
        # early loss read-back (ncf_set_loss_readback): train_step_host waits for the loss only, not for the whole step
        self._loss_host = torch.zeros(1).pin_memory() if os.environ.get("NCF_EARLY_LOSS", "1") != "0" else None
        self._loss_event = torch.cuda.Event() if self._loss_host is not None else None
        self._early_loss = False
        if max_rows:
            self._reserve(max_rows)

    def close(self):
        """Detach the auxiliary stream from the library (it must outlive every call that may use it)."""
        if getattr(self, "_aux_stream", None) is not None:
            torch.cuda.synchronize(self.device)
            with torch.cuda.device(self.device):
                self.lib.ncf_set_aux_stream(None)
            self._aux_stream = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    def _reserve(self, N):
        cfg = self._cfg()
        need = int(self.lib.ncf_workspace_bytes(N, C.byref(cfg)))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        if self._out is None or self._out.numel() < N:
            self._out = torch.empty(N, device=self.device)

    def _cfg(self):
        m = self.model
        cfg = _lib.RunCfg()
        cfg.S, cfg.training = self.S, 1
        cfg.precision = _lib.NCF_BF16_TC if m.compute_precision == "bf16" else _lib.NCF_FP32
        cfg.dropout_p = float(m.dropout)
        cfg.seed = m._dropout_seed
        cfg.step = self.step
        return cfg

    def _validate_model(self):
        """The full check (every dense parameter still a view of the flat buffer, tables contiguous fp32) walks all
        named parameters - too slow for every step of a loop that synchronises per batch.  It runs when the cheap
        signature (flat buffer, first / last dense view, the four table pointers) changes."""
        m = self.model
        tabs = m._table_params()
        flat = m._flat
        def signature():
            st = m._table_state or {}
            return ((m._flat.data_ptr(), self._dense_ends[0].data_ptr(), self._dense_ends[1].data_ptr())
                    + tuple(t.data_ptr() for t in tabs)
                    + tuple(t.data_ptr() for k in ("m", "v", "touched") for t in st.get(k, ())))
        if flat is None or self._dense_ends is None or signature() != self._model_sig:
            m._ensure_flat()
            dp = m._dense_params()
            self._dense_ends = (dp[0], dp[-1])
            self._tables = m._tables_struct()
            tabs = m._table_params()
            self._model_sig = signature()

    def train_step(self, user_ids: torch.Tensor, item_ids: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """ids int64 [N] and targets fp32 [N] on the device; returns the loss as a device scalar
        (no host sync).  Probabilities of this step stay in `self.outputs`."""
        N = user_ids.numel()
        if N % self.S:
            raise ValueError(f"{N} rows are not groups of {self.S}")
        self._validate_model()
        self.model.check_status("NCFTrainEngine.train_step (an earlier step)")
        self.step += 1
        self._reserve(N)
        cfg = self._cfg()
        adam = _lib.AdamCfg()
        adam.lr, (adam.beta1, adam.beta2) = self.hp["lr"], self.hp["betas"]
        adam.eps, adam.weight_decay, adam.step = self.hp["eps"], self.hp["weight_decay"], self.step
        adam.emb_mode = _lib.EMB_ADAM_DENSE_EQUIV if self.table_mode == "fused_dense_equiv" else _lib.EMB_ADAM_SPARSE
        tables = self._tables
        # the library keeps the auxiliary stream and its events per CURRENT device (ncf_set_aux_stream): make sure that is ours
        prev = torch.cuda.current_device()
        if prev != self.device.index:
            torch.cuda.set_device(self.device)
        try:
            self.lib.ncf_set_aux_stream(C.c_void_p(self._aux_stream.cuda_stream) if self._aux_stream is not None else None)
            if self._early_loss:
                self._loss_event.record(torch.cuda.current_stream(self.device))      # creates the event handle on this device
                self.lib.ncf_set_loss_readback(C.c_void_p(self._loss_host.data_ptr()), C.c_void_p(self._loss_event.cuda_event))
            else:
                self.lib.ncf_set_loss_readback(None, None)
            _lib.check(self.lib.ncf_train_step(C.byref(cfg), C.byref(adam), C.byref(tables), _lib.ptr(self.model._flat),
                                               _lib.ptr(self.dense_grad), _lib.ptr(self.dense_m), _lib.ptr(self.dense_v),
                                               _lib.ptr(user_ids), _lib.ptr(item_ids), _lib.ptr(targets), N,
                                               _lib.ptr(self._out), _lib.ptr(self.loss), _lib.ptr(self._ws),
                                               self._ws.numel(), _stream(self.device)), "ncf_train_step")
        finally:
            if prev != self.device.index:
                torch.cuda.set_device(prev)
        self.model._table_step = self.step
        self.outputs = self._out[:N]
        return self.loss

    def train_step_host(self, user_ids: torch.Tensor, item_ids: torch.Tensor, targets: torch.Tensor,
                        next_batch=None) -> float:
        """End-to-end step from HOST (ideally pinned) buffers: H2D copies of the ids and targets,
        the step, and the D2H read of the loss (what trainer.py:253-254, 289 do per batch).

        next_batch = the (user_ids, item_ids, targets) host tensors the NEXT call will pass: their H2D copies are
        queued on a copy stream into the second set of staging buffers while this step computes (plain input-pipeline
        double buffering; every step still copies its inputs and reads its loss)."""
        N = user_ids.numel()
        if self._dev_in is None or self._dev_in[0][0].numel() < N:
            self._dev_in = [(torch.empty(N, dtype=torch.long, device=self.device),
                             torch.empty(N, dtype=torch.long, device=self.device),
                             torch.empty(N, dtype=torch.float32, device=self.device)) for _ in range(2)]
            self._stage_slot, self._staged_key, self._staged_event = 0, None, None
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._slot_free = [None, None]          # event behind the last step that read each staging set
        cur = self._stage_slot
        du, di, dt = (b[:N] for b in self._dev_in[cur])
        key = (user_ids.data_ptr(), item_ids.data_ptr(), targets.data_ptr(), N)
        if self._staged_key == key:
            torch.cuda.current_stream(self.device).wait_event(self._staged_event)
        else:
            du.copy_(user_ids.reshape(-1), non_blocking=True)
            di.copy_(item_ids.reshape(-1), non_blocking=True)
            dt.copy_(targets.reshape(-1), non_blocking=True)
        self._early_loss = self._loss_host is not None
        try:
            loss = self.train_step(du, di, dt)
        finally:
            self._early_loss = False
        self._staged_key = None
        self._slot_free[cur] = torch.cuda.current_stream(self.device).record_event()
        if next_batch is not None:
            # the other staging set was last read by the PREVIOUS step, which may still be in its backward (the host runs
            # ahead since it waits for the loss only): the copies wait for the event behind that step
            nu, ni, nt = next_batch
            M = nu.numel()
            if M <= self._dev_in[cur ^ 1][0].numel():
                if self._slot_free[cur ^ 1] is not None:
                    self._copy_stream.wait_event(self._slot_free[cur ^ 1])
                with torch.cuda.stream(self._copy_stream):
                    for d, h in zip(self._dev_in[cur ^ 1], (nu, ni, nt)):
                        d[:M].copy_(h.reshape(-1), non_blocking=True)
                    self._staged_event = self._copy_stream.record_event()
                self._staged_key = (nu.data_ptr(), ni.data_ptr(), nt.data_ptr(), M)
        self._stage_slot = cur ^ 1
        if self._loss_host is not None:
            # the loss is final after the forward: wait for its copy only, the backward keeps running while the caller
            # prepares (and this engine enqueues) the next step
            self._loss_event.synchronize()
            value = float(self._loss_host[0])
        else:
            value = float(loss.item())
        self.model.check_status("NCFTrainEngine.train_step_host")      # K1 (which validates the ids) ran before the loss
        return value

    def state_dict(self):
        return {"step": self.step, "dense_m": self.dense_m.clone(), "dense_v": self.dense_v.clone(),
                "tables": self.model.table_optimizer_state_dict(), "hp": dict(self.hp)}

    def load_state_dict(self, sd):
        self.step = int(sd["step"])
        self.dense_m.copy_(sd["dense_m"])
        self.dense_v.copy_(sd["dense_v"])
        self.model.load_table_optimizer_state_dict(sd["tables"])


class ModelTrainer:
    """Mirror of reference ModelTrainer (trainer.py:27-95, 216-546) on the CUDA model."""

    def __init__(self, model: nn.Module, config: Dict[str, Any], num_gpus: int = 1,
                 table_mode: str = "fused_dense_equiv"):
        required = {"num_users", "num_products", "batch_size", "learning_rate"}          # trainer.py:36-61
        missing = required - set(config.keys())
        if missing:
            raise ValueError(f"Missing required parameters in config: {missing}")
        if not torch.cuda.is_available():
            raise _lib.NcfError("ModelTrainer needs a CUDA device (no CPU fallback)")
        self.model = model
        self.config = config
        self.device = torch.device("cuda")
        self.num_gpus = num_gpus
        self.negative_samples = config.get("negative_samples", 4)
        self.logger = logging.getLogger(__name__)
        weight_decay = float(config.get("weight_decay", 1e-5))
        self.model = self.model.to(self.device)
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=config["learning_rate"],
                                          weight_decay=weight_decay)                        # trainer.py:71-75
        self.criterion = nn.BCELoss().to(self.device)                                       # trainer.py:78
        if isinstance(self.model, AdvancedNCF):
            self.model.configure_table_optimizer(table_mode, optimizer=self.optimizer)

    def train_epoch(self, train_loader) -> float:
        """trainer.py:216-337."""
        self.model.train()
        total_loss, num_batches = 0.0, 0
        for batch_idx, (features, targets) in enumerate(train_loader):
            base = len(features.lengths()) // len(features.keys())
            if base < 2:                                                                    # trainer.py:248-250
                logging.warning(f"Skipping small batch {batch_idx}: size {base}")
                continue
            features = features.to(self.device)
            targets = targets.to(self.device)
            outputs = self.model(features)
            if outputs.shape != targets.shape:
                targets = targets.view(outputs.shape)
            loss = self.criterion(outputs, targets)
            self.optimizer.zero_grad()
            loss.backward()
            self.optimizer.step()
            total_loss += loss.item()
            num_batches += 1
        avg = total_loss / num_batches if num_batches > 0 else float("inf")
        logging.info(f"Epoch complete - Average loss: {avg:.4f}")
        return avg

    def validate(self, val_loader, negative_samples: int = 0) -> Dict[str, float]:
        """trainer.py:350-410 (metrics stay on the device until the final scalars)."""
        self.model.eval()
        total_loss, outs, tgts, nb = 0.0, [], [], 0
        with torch.no_grad():
            for features, targets in val_loader:
                features = features.to(self.device)
                targets = targets.to(self.device)
                outputs = self.model(features)
                total_loss += self.criterion(outputs, targets.view(outputs.shape)).item()
                outs.append(outputs)
                tgts.append(targets.view(outputs.shape))
                nb += 1
        all_o, all_t = torch.cat(outs, 0), torch.cat(tgts, 0)
        groups = all_t.numel() // (1 + negative_samples)
        metrics = calculate_metrics(all_o, all_t, [1, 5, 10], batch_size=groups, negative_samples=negative_samples)
        metrics["loss"] = total_loss / max(nb, 1)
        return metrics

    def train(self, train_loader, val_loader, num_epochs: int, early_stopping_patience: int = 5,
              checkpoint_dir: Optional[str] = None) -> Dict[str, List[float]]:
        """trainer.py:412-546: epochs, early stopping on validation loss, best/latest checkpoints."""
        history: Dict[str, List[float]] = {"train_loss": [], "val_loss": []}
        best, bad = float("inf"), 0
        start = 0
        if checkpoint_dir:
            os.makedirs(checkpoint_dir, exist_ok=True)
            latest = self._find_latest_checkpoint(checkpoint_dir)
            if latest:
                start = self._load_checkpoint(latest)
        for epoch in range(start, num_epochs):
            tl = self.train_epoch(train_loader)
            vm = self.validate(val_loader)
            history["train_loss"].append(tl)
            history["val_loss"].append(vm["loss"])
            for k, v in vm.items():
                history.setdefault(f"val_{k}", []).append(v)
            improved = vm["loss"] < best
            if improved:
                best, bad = vm["loss"], 0
            else:
                bad += 1
            if checkpoint_dir:
                self._save_checkpoint(checkpoint_dir, epoch, vm, is_best=improved)
            if bad >= early_stopping_patience:
                logging.info(f"Early stopping after {epoch + 1} epochs")
                break
        return history

    # ---- checkpoints: the reference's format (trainer.py:548-609) ------------------------------------------------
    def _table_param_indices(self):
        """positions of the four table parameters in `model.parameters()` order = their keys in optimizer.state_dict()"""
        if not isinstance(self.model, AdvancedNCF):
            return []
        ids = [id(t) for t in self.model._table_params()]
        pos = {id(p): k for k, p in enumerate(self.model.parameters())}
        return [pos[i] for i in ids]

    def _optimizer_state_with_tables(self):
        """optimizer.state_dict() as the REFERENCE would have written it: in the fused table modes the tables' Adam moments
        live in the module (their .grad stays None, so torch's Adam holds no state for them); mirror them into the
        entries torch.optim.Adam uses (step / exp_avg / exp_avg_sq), so a reference-side resume finds them."""
        sd = self.optimizer.state_dict()
        m = self.model
        if isinstance(m, AdvancedNCF) and m._table_state is not None:
            sd = {"state": dict(sd["state"]), "param_groups": sd["param_groups"]}
            for k, idx in enumerate(self._table_param_indices()):
                sd["state"][idx] = {"step": torch.tensor(float(m._table_step)), "exp_avg": m._table_state["m"][k].clone(),
                                    "exp_avg_sq": m._table_state["v"][k].clone()}
        return sd

    def _save_checkpoint(self, checkpoint_dir, epoch, metrics, is_best=False, filename=None):
        m = self.model
        if filename is None:
            filename = f"checkpoint_epoch_{epoch + 1}.pt"                                   # trainer.py:558
        ck = {"epoch": epoch, "model_state_dict": m.state_dict(), "optimizer_state_dict": self._optimizer_state_with_tables(),
              "metrics": metrics, "config": self.config,
              "model_config": {"num_users": m.num_users, "num_products": m.num_products,
                               "embedding_dim": m.mf_embedding_dim}}                       # trainer.py:566-570
        path = os.path.join(checkpoint_dir, filename)
        torch.save(ck, path)
        if is_best:
            best = os.path.join(checkpoint_dir, "best_model.pt")
            if os.path.lexists(best):
                os.remove(best)
            os.symlink(filename, best)
        return path

    @staticmethod
    def _find_latest_checkpoint(checkpoint_dir):
        """the helper the reference calls but never defines (trainer.py:450)."""
        best, best_e = None, -1
        for f in os.listdir(checkpoint_dir):
            if f.startswith("checkpoint_epoch_") and f.endswith(".pt"):
                try:
                    e = int(f[len("checkpoint_epoch_"):-3])
                except ValueError:
                    continue
                if e > best_e:
                    best, best_e = os.path.join(checkpoint_dir, f), e
        return best

    def _load_checkpoint(self, path) -> int:
        """trainer.py:588-609: returns the epoch to resume FROM (saved epoch + 1).  A checkpoint written by the reference
        (or by the "autograd" table mode) keeps the tables' moments inside optimizer_state_dict: in the fused modes they
        are moved into the module's table state, so the resumed run continues the same Adam trajectory."""
        ck = torch.load(path, map_location=self.device, weights_only=False)
        self.model.load_state_dict(ck["model_state_dict"])
        osd = ck["optimizer_state_dict"]
        m = self.model
        fused = isinstance(m, AdvancedNCF) and m._table_mode != "autograd"
        if fused:
            osd = {"state": dict(osd["state"]), "param_groups": osd["param_groups"]}
            idxs = self._table_param_indices()
            if all(i in osd["state"] for i in idxs):
                st = [osd["state"].pop(i) for i in idxs]
                m.load_table_optimizer_state_dict({"step": int(float(st[0]["step"])), "m": [x["exp_avg"] for x in st],
                                                   "v": [x["exp_avg_sq"] for x in st]})
            elif ck.get("table_optimizer_state"):                                          # round-1 format of this repo
                m.load_table_optimizer_state_dict(ck["table_optimizer_state"])
        self.optimizer.load_state_dict(osd)
        return int(ck["epoch"]) + 1

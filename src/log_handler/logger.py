"""Drop-in for src/log_handler/logger.py:8-18 (TensorBoard scalars; silently a no-op without tensorboard)."""


class INRLogger:
    def __init__(self, log_dir):
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.w = SummaryWriter(log_dir)
        except Exception:
            self.w = None

    def log_train(self, loss, it):
        if self.w:
            self.w.add_scalar("train_loss", loss, it)

    def log_test(self, loss, psnr, ssim, epoch):
        if self.w:
            self.w.add_scalar("test_loss", loss, epoch)
            self.w.add_scalar("test_psnr", psnr, epoch)
            self.w.add_scalar("test_ssim", ssim, epoch)

"""Drop-in check: the reference's LightningModule body (movenet/pytorch_lightning_trainer.py:24-74) replayed line by
line on top of ``from movenet.wavenet import WaveNet`` -- the import the trainer itself uses, resolved by the shim
package ``movenet/`` of this repository.  pytorch_lightning is not installed here, so the class below restates
``Dance2Music`` without the Lightning base class; every line that touches the model is the reference's."""
from dataclasses import asdict, dataclass

import pytest
import torch
import torch.nn.functional as F

from conftest import golden_audio, golden_video, load_golden

pytestmark = pytest.mark.gpu


@dataclass
class ModelConfig:                      # movenet/config.py:11-18
    layer_size: int = 3
    stack_size: int = 3
    input_channels: int = 256
    residual_channels: int = 16
    skip_channels: int = 16


class Dance2Music(torch.nn.Module):     # movenet/pytorch_lightning_trainer.py:24-74 minus Lightning
    def __init__(self, model_config, use_video=True, generate_n_samples=None, generate_temperature=0.0):
        super().__init__()
        from movenet.wavenet import WaveNet                       # :16
        self.model = WaveNet(**asdict(model_config))              # :31
        self.use_video, self.precision = use_video, 32
        self.generate_n_samples, self.generate_temperature = generate_n_samples, generate_temperature
        self.logged = {}

    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, audio, video, **kwargs):
        return self.model(audio, video, **kwargs)                 # :33-34

    def generate(self, audio, video):
        return self.model.generate(audio, video, n_samples=self.generate_n_samples,
                                   temperature=self.generate_temperature).detach()      # :43-49

    def training_step(self, batch, batch_idx):
        audio, video, contexts, fps, info = batch
        dtype = getattr(torch, f"float{self.precision}")
        audio = audio.type(dtype).to(self.device)
        if self.use_video:
            video = video.type(dtype).to(self.device)
        output = self(audio, video)
        target = audio[:, :, self.model.receptive_fields:].argmax(1)
        loss = F.cross_entropy(output, target)
        acc = (output.argmax(1) == target).float().mean()
        self.logged.update(train_loss=loss, train_acc=acc)
        return {"loss": loss, "output": output.detach(), "generated_output": self.generate(audio, video)}


def test_lightning_training_step_runs_unchanged_through_the_shim():
    import movenet_b200
    fx = load_golden("video")
    cfg = ModelConfig(**fx["shape"])
    RF = movenet_b200.WaveNet(**fx["shape"]).receptive_fields
    module = Dance2Music(cfg, use_video=True, generate_n_samples=RF + 12, generate_temperature=0.0)
    assert type(module.model) is movenet_b200.WaveNet
    module.model.load_state_dict(fx["params"], strict=False)
    module.cuda()
    opt = torch.optim.AdamW(module.parameters(), lr=3e-4)         # :128-202 default optimizer
    audio, video = golden_audio(fx), golden_video(fx, 1)          # host tensors, like a DataLoader batch
    losses = []
    for step in range(2):
        opt.zero_grad()
        out = module.training_step((audio, video, None, None, None), step)
        out["loss"].backward()
        opt.step()
        losses.append(out["loss"].item())
        assert out["output"].shape == (1, cfg.input_channels, 160000 - RF)
        gen = out["generated_output"]
        assert gen.shape == (1, cfg.input_channels, RF + 12) and torch.equal(gen.sum(1), torch.ones_like(gen.sum(1)))
        assert torch.equal(gen[:, :, :RF].cpu(), audio[:, :, :RF])
    assert abs(losses[0] - fx["loss"].item()) <= 1e-3 * abs(fx["loss"].item())
    assert losses[1] != losses[0] and all(map(lambda v: v == v, losses))
    assert not module.model.training      # generate() leaves the model in eval mode, like the reference (wavenet.py:202)

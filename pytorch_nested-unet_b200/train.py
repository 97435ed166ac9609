"""Drop-in for the reference's training script (trains.py; README calls it train.py): same flags and defaults
(trains.py:31-103), same per-batch body (trains.py:113-136 / 159-177), same artefacts (models/<name>/config.yml,
log.csv with epoch, lr, loss, iou, val_loss, val_iou, best-IoU model.pth), running on the B200 engine.

Data: the reference's loader needs cv2 + albumentations and the DSB2018 files (out of this build's scope, SURVEY.md 2
rows 8-9); `--dataset synthetic` (the default when inputs/<dataset> does not exist) draws batches with the loader's
output distribution: x = ((u8/255 - mean)/std)/255 (albumentations Normalize() then dataset.py:71), mask in {0,1}.
Multi-GPU: launch with torchrun; the model is wrapped in dist.DataParallel (SyncBN + overlapped gradient all-reduce).
"""
import argparse
import os
import sys
from collections import OrderedDict

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import torch
import optim
import yaml
from torch.optim import lr_scheduler

import archs
import losses
from metrics import iou_counts_device, iou_score
from utils import AverageMeter, DevicePrefetcher, StepScalars, device_target, settle_gc, str2bool

ARCH_NAMES = archs.__all__
LOSS_NAMES = losses.__all__ + ['BCEWithLogitsLoss']      # trains.py:27-28


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--name', default=None, help='model name: (default: arch+timestamp)')
    parser.add_argument('--epochs', default=100, type=int, metavar='N', help='number of total epochs to run')
    parser.add_argument('-b', '--batch_size', default=16, type=int, metavar='N', help='mini-batch size (default: 16)')
    # model
    parser.add_argument('--arch', '-a', metavar='ARCH', default='NestedUNet', choices=ARCH_NAMES,
                        help='model architecture: ' + ' | '.join(ARCH_NAMES) + ' (default: NestedUNet)')
    parser.add_argument('--deep_supervision', default=False, type=str2bool)
    parser.add_argument('--input_channels', default=3, type=int, help='input channels')
    parser.add_argument('--num_classes', default=1, type=int, help='number of classes')
    parser.add_argument('--input_w', default=96, type=int, help='image width')
    parser.add_argument('--input_h', default=96, type=int, help='image height')
    # loss
    parser.add_argument('--loss', default='BCEDiceLoss', choices=LOSS_NAMES,
                        help='loss: ' + ' | '.join(LOSS_NAMES) + ' (default: BCEDiceLoss)')
    # dataset
    parser.add_argument('--dataset', default='dsb2018_96', help='dataset name')
    parser.add_argument('--img_ext', default='.png', help='image file extension')
    parser.add_argument('--mask_ext', default='.png', help='mask file extension')
    # optimizer
    parser.add_argument('--optimizer', default='SGD', choices=['Adam', 'SGD'],
                        help='loss: ' + ' | '.join(['Adam', 'SGD']) + ' (default: Adam)')
    parser.add_argument('--lr', '--learning_rate', default=1e-3, type=float, metavar='LR', help='initial learning rate')
    parser.add_argument('--momentum', default=0.9, type=float, help='momentum')
    parser.add_argument('--weight_decay', default=1e-4, type=float, help='weight decay')
    parser.add_argument('--nesterov', default=False, type=str2bool, help='nesterov')
    # scheduler
    parser.add_argument('--scheduler', default='CosineAnnealingLR',
                        choices=['CosineAnnealingLR', 'ReduceLROnPlateau', 'MultiStepLR', 'ConstantLR'])
    parser.add_argument('--min_lr', default=1e-5, type=float, help='minimum learning rate')
    parser.add_argument('--factor', default=0.1, type=float)
    parser.add_argument('--patience', default=2, type=int)
    parser.add_argument('--milestones', default='1,2', type=str)
    parser.add_argument('--gamma', default=2 / 3, type=float)
    parser.add_argument('--early_stopping', default=-1, type=int, metavar='N', help='early stopping (default: -1)')
    parser.add_argument('--num_workers', default=4, type=int)
    # synthetic-data extension (not in the reference)
    parser.add_argument('--synthetic_batches', default=8, type=int, help='batches per epoch when --dataset synthetic')
    parser.add_argument('--u8_input', default=True, type=str2bool,
                        help='synthetic loader yields raw uint8 NHWC images / masks; Normalize() + /255 run on the GPU (4x fewer H2D bytes)')
    return parser.parse_args(argv)


def _world_size():
    import torch.distributed as dist
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def _rank():
    import torch.distributed as dist
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def _iou(output, target):
    """metrics.iou_score (metrics.py:6-18); under data parallelism over the GLOBAL batch (integer counts are all-reduced)."""
    if _world_size() > 1:
        import dist as nunet_dist
        return nunet_dist.global_iou_score(output, target)
    return iou_score(output, target)


def _global_mean(value, device):
    if _world_size() > 1:
        import dist as nunet_dist
        return nunet_dist.global_mean(value, device)
    return value


def _loss_and_iou(config, model, criterion, input, target, on_device=False):
    """trains.py:117-127.  on_device: the IoU is returned as its two integer counts on the device (no host synchronisation)."""
    iou_fn = iou_counts_device if on_device else _iou
    if config['deep_supervision']:
        outputs = model(input)
        loss = 0
        for output in outputs:
            loss += criterion(output, target)
        loss /= len(outputs)
        iou = iou_fn(outputs[-1], target)
    else:
        output = model(input)
        loss = criterion(output, target)
        iou = iou_fn(output, target)
    return loss, iou


def train(config, train_loader, model, criterion, optimizer):
    """trains.py:106-147.  On one GPU the scalars of a step (trains.py:124,133-134) are read through utils.StepScalars — their
    device->host copy is issued before backward is enqueued and waited for after the optimizer step has been enqueued, so the CPU
    never waits for more than the forward pass and the GPU never runs dry — and batches arrive through utils.DevicePrefetcher (copy of
    batch i+1 on a copy stream while step i runs); the meters see the same values in the same iteration.
    Under data parallelism the loop keeps the reference's placement (copy on the compute stream, scalars read after the step): with
    the overlapped placement loss.backward() occasionally blocked the host for 30-190 ms on both ranks (2 GPUs, 5 of 11 twenty-step
    runs; not root-caused), which costs more than the 3 % the overlap gains."""
    avg_meters = {'loss': AverageMeter(), 'iou': AverageMeter()}
    model.train()
    overlap = _world_size() == 1
    scalars = StepScalars(reduce=True)
    steps_done = 0
    for input, target, _ in (DevicePrefetcher(train_loader) if overlap else train_loader):
        input = input.cuda(non_blocking=True)
        target = device_target(target.cuda(non_blocking=True))
        loss, counts = _loss_and_iou(config, model, criterion, input, target, on_device=True)
        if overlap:
            scalars.submit(loss, counts)
            scalars.before_backward()
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        if not overlap:
            scalars.submit(loss, counts)       # global mean loss / summed counts on the device, then one host read
        for lv, i_, u_ in scalars.collect():
            avg_meters['loss'].update(lv, input.size(0))
            avg_meters['iou'].update((i_ + 1e-5) / (u_ + 1e-5), input.size(0))      # metrics.py:17-18
        steps_done += 1
        if steps_done == 3:
            settle_gc()        # after the eager, the capturing and the first replayed step: the heap is static from here on
    return OrderedDict([('loss', avg_meters['loss'].avg), ('iou', avg_meters['iou'].avg)])


def validate(config, val_loader, model, criterion):
    """trains.py:150-188."""
    avg_meters = {'loss': AverageMeter(), 'iou': AverageMeter()}
    model.eval()
    with torch.no_grad():
        for input, target, _ in val_loader:
            input = input.cuda(non_blocking=True)
            target = device_target(target.cuda(non_blocking=True))
            loss, iou = _loss_and_iou(config, model, criterion, input, target)
            avg_meters['loss'].update(_global_mean(loss.item(), input.device), input.size(0))
            avg_meters['iou'].update(iou, input.size(0))
    return OrderedDict([('loss', avg_meters['loss'].avg), ('iou', avg_meters['iou'].avg)])


class SyntheticLoader(object):
    """Yields (img, mask, meta) batches with the shape/distribution of dataset.py:55-76 after trains.py:266."""

    def __init__(self, batches, batch_size, channels, classes, h, w, seed, u8=False):
        self.batches, self.shape, self.classes, self.seed, self.u8 = batches, (batch_size, channels, h, w), classes, seed, u8

    def __len__(self):
        return self.batches

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        n, c, h, w = self.shape
        mean = torch.tensor([(0.485, 0.456, 0.406)[i % 3] for i in range(c)]).view(1, c, 1, 1)
        std = torch.tensor([(0.229, 0.224, 0.225)[i % 3] for i in range(c)]).view(1, c, 1, 1)
        for b in range(self.batches):
            raw = torch.randint(0, 256, (n, c, h, w), generator=g, dtype=torch.uint8)
            u8 = raw.float()
            # blob-like masks correlated with the image so that the IoU can actually improve
            t = (torch.nn.functional.avg_pool2d(u8[:, :1], 9, 1, 4) > 128).float().repeat(1, self.classes, 1, 1)
            meta = {'img_id': ['syn%d_%d' % (b, i) for i in range(n)]}
            if self.u8:   # what cv2.imread yields (HWC uint8, dataset.py:58-68): Normalize() + /255 then run on the GPU
                yield raw.permute(0, 2, 3, 1).contiguous().pin_memory(), (t * 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory(), meta
            else:
                yield (((u8 / 255.0 - mean) / std) / 255.0).pin_memory(), t.pin_memory(), meta


def main(argv=None):
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per engine stream (see bench.py)
    config = vars(parse_args(argv))
    if config['name'] is None:
        config['name'] = '%s_%s_%s' % (config['dataset'], config['arch'], 'wDS' if config['deep_supervision'] else 'woDS')
    # multi-GPU (torchrun): every rank trains on its own shard of the data; rank 0 alone writes config.yml / log.csv / model.pth
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if rank == 0:
        os.makedirs('models/%s' % config['name'], exist_ok=True)
        print('-' * 20)
        for key in config:
            print('%s: %s' % (key, config[key]))
        print('-' * 20)
        with open('models/%s/config.yml' % config['name'], 'w') as f:
            yaml.dump(config, f)

    if config['loss'] == 'BCEWithLogitsLoss':
        criterion = torch.nn.BCEWithLogitsLoss().cuda()
    else:
        criterion = losses.__dict__[config['loss']]().cuda()

    if rank == 0:
        print("=> creating model %s" % config['arch'])
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
            dist.init_process_group('nccl')
    model = archs.__dict__[config['arch']](config['num_classes'], config['input_channels'], config['deep_supervision'])
    model = model.cuda()
    if world > 1:
        import dist as nunet_dist
        model = nunet_dist.DataParallel(model)

    params = filter(lambda p: p.requires_grad, model.parameters())
    # same constructors as trains.py:225-233; `optim` here is this package's drop-in (torch.optim subclasses whose step() is one
    # fused multi-tensor launch, optim.py)
    if config['optimizer'] == 'Adam':
        optimizer = optim.Adam(params, lr=config['lr'], weight_decay=config['weight_decay'])
    else:
        optimizer = optim.SGD(params, lr=config['lr'], momentum=config['momentum'], nesterov=config['nesterov'],
                              weight_decay=config['weight_decay'])
    if config['scheduler'] == 'CosineAnnealingLR':
        scheduler = lr_scheduler.CosineAnnealingLR(optimizer, T_max=config['epochs'], eta_min=config['min_lr'])
    elif config['scheduler'] == 'ReduceLROnPlateau':
        scheduler = lr_scheduler.ReduceLROnPlateau(optimizer, factor=config['factor'], patience=config['patience'], min_lr=config['min_lr'])
    elif config['scheduler'] == 'MultiStepLR':
        scheduler = lr_scheduler.MultiStepLR(optimizer, milestones=[int(e) for e in config['milestones'].split(',')], gamma=config['gamma'])
    else:
        scheduler = None

    data_dir = os.path.join('inputs', config['dataset'])
    if config['dataset'] != 'synthetic' and os.path.isdir(data_dir):
        raise NotImplementedError("the cv2/albumentations file loader of dataset.py is outside this build's scope; use --dataset synthetic")
    mk = lambda seed, nb: SyntheticLoader(nb, config['batch_size'], config['input_channels'], config['num_classes'],
                                          config['input_h'], config['input_w'], seed, u8=config.get('u8_input', False))
    # each rank draws its own batches (seed offset by the rank): the global batch is world x batch_size distinct images
    train_loader, val_loader = mk(41 + 1000 * rank, config['synthetic_batches']), mk(42 + 1000 * rank, max(1, config['synthetic_batches'] // 4))

    log = OrderedDict([('epoch', []), ('lr', []), ('loss', []), ('iou', []), ('val_loss', []), ('val_iou', [])])
    best_iou, trigger = 0, 0
    for epoch in range(config['epochs']):
        if rank == 0:
            print('Epoch [%d/%d]' % (epoch, config['epochs']))
        train_log = train(config, train_loader, model, criterion, optimizer)
        val_log = validate(config, val_loader, model, criterion)
        if config['scheduler'] == 'CosineAnnealingLR' or config['scheduler'] == 'MultiStepLR':
            scheduler.step()
        elif config['scheduler'] == 'ReduceLROnPlateau':
            scheduler.step(val_log['loss'])
        if rank != 0:       # metrics above are already global (all-reduced); only rank 0 logs and checkpoints
            trigger += 1
            if val_log['iou'] > best_iou:
                best_iou, trigger = val_log['iou'], 0
            if config['early_stopping'] >= 0 and trigger >= config['early_stopping']:
                break
            torch.cuda.empty_cache()
            continue
        print('loss %.4f - iou %.4f - val_loss %.4f - val_iou %.4f' % (train_log['loss'], train_log['iou'], val_log['loss'], val_log['iou']))
        log['epoch'].append(epoch)
        log['lr'].append(config['lr'])
        log['loss'].append(train_log['loss'])
        log['iou'].append(train_log['iou'])
        log['val_loss'].append(val_log['loss'])
        log['val_iou'].append(val_log['iou'])
        with open('models/%s/log.csv' % config['name'], 'w') as f:
            f.write(','.join([''] + list(log.keys())) + '\n')
            for i in range(len(log['epoch'])):
                f.write(','.join([str(i)] + [str(log[k][i]) for k in log]) + '\n')
        trigger += 1
        if val_log['iou'] > best_iou:
            torch.save(model.state_dict(), 'models/%s/model.pth' % config['name'])
            best_iou = val_log['iou']
            print("=> saved best model")
            trigger = 0
        if config['early_stopping'] >= 0 and trigger >= config['early_stopping']:
            print("=> early stopping")
            break
        torch.cuda.empty_cache()
    return log


if __name__ == '__main__':
    main()

"""Drop-in for the reference's val.py (val.py:31-109): reloads models/<name>/config.yml and model.pth, runs the
eval-mode forward (BatchNorm folded to a per-channel affine inside the conv loaders), prints the batch-size weighted
mean of iou_score.  The per-image output stage (val.py:100-105) runs sigmoid, x255 and the uint8 cast on the device
(ops.sigmoid_u8) and writes outputs/<name>/<class>/<img_id>.jpg with cv2 as the reference does; the file loader is synthetic for the
same reason as in train.py."""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import numpy as np
import torch
import yaml

import archs
from metrics import iou_score
from utils import AverageMeter


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--name', default=None, help='model name')
    parser.add_argument('--no_dump', action='store_true', help='skip writing outputs/<name>/<class>/<img_id>.jpg (val.py:102-105)')
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    with open('models/%s/config.yml' % args.name, 'r') as f:
        config = yaml.load(f, Loader=yaml.FullLoader)
    print('-' * 20)
    for key in config.keys():
        print('%s: %s' % (key, str(config[key])))
    print('-' * 20)
    print("=> creating model %s" % config['arch'])
    model = archs.__dict__[config['arch']](config['num_classes'], config['input_channels'], config['deep_supervision'])
    model = model.cuda()
    model.load_state_dict(torch.load('models/%s/model.pth' % config['name']))
    model.eval()
    from train import SyntheticLoader
    from utils import device_target
    import ops
    val_loader = SyntheticLoader(max(1, config.get('synthetic_batches', 8) // 4), config['batch_size'], config['input_channels'],
                                 config['num_classes'], config['input_h'], config['input_w'], 42, u8=config.get('u8_input', False))
    try:
        import cv2
    except ImportError:     # the reference writes JPEGs with cv2 (val.py:102-105); without it the planes are kept as .npy
        cv2 = None
    avg_meter = AverageMeter()
    for c in range(config['num_classes']):
        os.makedirs(os.path.join('outputs', config['name'], str(c)), exist_ok=True)
    with torch.no_grad():
        for input, target, meta in val_loader:
            input = input.cuda()
            target = device_target(target.cuda())
            if config['deep_supervision']:
                output = model(input)[-1]
            else:
                output = model(input)
            iou = iou_score(output, target)
            avg_meter.update(iou, input.size(0))
            if not args.no_dump:
                # val.py:100-105: sigmoid, x255 and the uint8 cast on the device; only the uint8 planes cross to the host
                prob = ops.sigmoid_u8(output).cpu().numpy()
                for i in range(len(prob)):
                    for c in range(config['num_classes']):
                        path = os.path.join('outputs', config['name'], str(c), meta['img_id'][i])
                        if cv2 is not None:
                            cv2.imwrite(path + '.jpg', prob[i, c])
                        else:
                            np.save(path + '.npy', prob[i, c])
    print('IoU: %.4f' % avg_meter.avg)
    torch.cuda.empty_cache()
    return avg_meter.avg


if __name__ == '__main__':
    main()

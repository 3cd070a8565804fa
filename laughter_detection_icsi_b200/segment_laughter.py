"""Drop-in for the reference's segment_laughter.py CLI (:28-52 flags, :59-74 model loading, :79-122 load_and_pred,
:124-161 save_instances), running features + ResNetBigger + run extraction on one B200.

    python -m laughter_detection_icsi_b200.segment_laughter --config=resnet_base --model_path=<dir with best.pth.tar> \
        --input_audio_file=chan0.wav --thresholds=0.2,0.5 --min_lengths=0.1,0.2 --save_to_textgrid=True \
        --save_to_audio_files=False --output_dir=out

Output tree: <output_dir>/t_<thr>/l_<min_len>/<audio basename>.TextGrid (nothing is written for a setting
without instances, as in the reference).
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.io.wavfile
import torch

from . import config as config_module
from . import laugh_segmenter, load_data, textgrid
from .utils import audio_utils, torch_utils


def strtobool(val):
    val = str(val).lower()
    if val in ("y", "yes", "t", "true", "on", "1"):
        return 1
    if val in ("n", "no", "f", "false", "off", "0"):
        return 0
    raise ValueError("invalid truth value %r" % (val,))


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--model_path', type=str, default='checkpoints/in_use/resnet_with_augmentation')
    parser.add_argument('--config', type=str, default='resnet_with_augmentation')
    parser.add_argument('--thresholds', '--threshold', dest='thresholds', type=str, default='0.5',
                        help='Single value or comma-separated list of thresholds to evaluate')
    parser.add_argument('--min_lengths', '--min_length', dest='min_lengths', type=str, default='0.2',
                        help='Single value or comma-separated list of min_lengths to evaluate')
    parser.add_argument('--input_audio_file', required=True, type=str)
    parser.add_argument('--output_dir', type=str, default=None)
    parser.add_argument('--save_to_audio_files', type=str, default='True')
    parser.add_argument('--save_to_textgrid', type=str, default='False')
    return parser


def load_model(model_path, config, device):
    """ResNetBigger(dropout_rate=0.0, ...) on `device` with <model_path>/best.pth.tar loaded, in eval mode."""
    model = config['model'](dropout_rate=0.0, linear_layer_size=config['linear_layer_size'],
                            filter_sizes=config['filter_sizes'])
    model.set_device(device)
    if os.path.exists(model_path):
        torch_utils.load_checkpoint(model_path + '/best.pth.tar', model)
        model.eval()
    else:
        raise Exception(f"Model checkpoint not found at {model_path}")
    return model


def predict_probs(model, audio_path):
    """One probability per 10 ms frame: fused features -> network on the GPU."""
    return load_data.infer_audio_file(audio_path, model)


def load_and_pred(audio_path, model, thresholds, min_lengths, output_dir, save_to_audio_files, save_to_textgrid):
    """Predicts, segments and writes outputs; returns the time taken excluding output files."""
    start_time = time.time()
    probs = load_data.infer_audio_file(audio_path, model, as_tensor=True)   # stays on the model's GPU for the segmenter
    file_length = audio_utils.get_audio_length(audio_path)
    fps = len(probs) / float(file_length)
    # the reference disabled laugh_segmenter.lowpass here because it can output probs < 0 (segment_laughter.py:107-108)
    instance_dict = laugh_segmenter.get_laughter_instances(probs, thresholds=thresholds, min_lengths=min_lengths, fps=fps)
    time_taken = time.time() - start_time
    print(f'Completed in: {time_taken:.2f}s')
    for setting, instances in instance_dict.items():
        print(f"Found {len(instances)} laughs for threshold {setting[0]} and min_length {setting[1]}.")
        instance_output_dir = os.path.join(output_dir, f't_{setting[0]}', f'l_{setting[1]}')
        save_instances(instances, instance_output_dir, save_to_audio_files, save_to_textgrid, audio_path)
    return time_taken


def save_instances(instances, output_dir, save_to_audio_files, save_to_textgrid, audio_path):
    os.makedirs(output_dir, exist_ok=True)
    if len(instances) == 0:
        return
    if save_to_audio_files:
        # The reference reloads the file at 44.1 kHz with librosa; here laughs are cut at the file's own rate.
        y, sr = audio_utils.load_wav_int16(audio_path)
        wav_paths = []
        for index, instance in enumerate(instances):
            laughs = laugh_segmenter.cut_laughter_segments([instance], y, sr)
            wav_path = output_dir + "/laugh_" + str(index) + ".wav"
            scipy.io.wavfile.write(wav_path, sr, np.asarray(laughs).astype(np.int16))
            wav_paths.append(wav_path)
        print(laugh_segmenter.format_outputs(instances, wav_paths))
    if save_to_textgrid:
        fname = os.path.splitext(os.path.basename(audio_path))[0]
        path = os.path.join(output_dir, fname + '.TextGrid')
        textgrid.write_laughter_textgrid(path, instances)
        print('Saved laughter segments in {}'.format(path))


def calc_real_time_factor(audio_path, iterations, model, **kw):
    if not os.path.isfile(audio_path):
        raise ValueError(f"Audio_path doesn't exist. Given path {audio_path}")
    audio_length = audio_utils.get_audio_length(audio_path)
    print(f"Audio Length: {audio_length}")
    total = sum(load_and_pred(audio_path, model, **kw) for _ in range(iterations))
    rtf = total / iterations / audio_length
    print(f"Average Realtime Factor over {iterations} iterations: {rtf:.2f}")
    return rtf


def main(argv=None):
    args = build_parser().parse_args(argv)
    config = config_module.MODEL_MAP[args.config]
    save_to_audio_files = bool(strtobool(args.save_to_audio_files))
    save_to_textgrid = bool(strtobool(args.save_to_textgrid))
    thresholds = [float(t) for t in args.thresholds.split(',')]
    min_lengths = [float(l) for l in args.min_lengths.split(',')]
    if not torch.cuda.is_available():
        raise Exception("No CUDA device found: this build runs on B200 only (no CPU path)")
    device = torch.device('cuda')
    print(f"Using device {device}")
    if args.output_dir is None:
        raise Exception("Need to specify an output directory")
    model = load_model(args.model_path, config, device)
    return load_and_pred(args.input_audio_file, model, thresholds, min_lengths, args.output_dir, save_to_audio_files,
                         save_to_textgrid)


if __name__ == '__main__':
    main()
    sys.exit(0)

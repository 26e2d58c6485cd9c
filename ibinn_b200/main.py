"""`python -m ibinn_b200.main [train|test] <config file>`: the reference's entry point (reference main.py:1-38):
default.ini overlaid by the experiment .ini, output folder + conf.ini dump, dispatch."""
import os
import sys

from . import config

USAGE = "Usage: main.py [train|test] <config file>"


def main(argv=None):
    argv = sys.argv if argv is None else argv
    assert len(argv) in [2, 3], USAGE
    mode = argv[1]
    conf_file = argv[2] if len(argv) == 3 else None
    if conf_file is not None:
        assert os.path.isfile(conf_file), "No such config file"
    args = config.make_args(ini=conf_file)
    output_dir = os.path.join(args['checkpoints']['global_output_folder'], args['checkpoints']['base_name'])
    args['checkpoints']['output_dir'] = output_dir
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, 'conf.ini'), 'w') as f:
        args.write(f)
    if mode == 'train':
        from . import train
        train.train(args, synthetic=os.environ.get('IBINN_SYNTHETIC_DATA', '0') == '1')   # opt-in only; otherwise the reference's data.py
    elif mode == 'test':
        raise SystemExit('`test` runs the reference\'s evaluation/ package (plots, OoD, calibration), which is outside the '
                         'ibinn_b200 hot path: point it at ibinn_b200.model.GenerativeClassifier (see INTEGRATION.md)')
    else:
        print(USAGE)
        sys.exit(1)


if __name__ == '__main__':
    main()

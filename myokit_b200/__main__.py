"""
Command line: ``python -m myokit_b200 cuda`` and ``cuda-select``, the
counterparts of ``myokit opencl`` / ``myokit opencl-select``
(``myokit/__main__.py:738-880``). A CUDA device is an ordinal, so selection is
one number (``--device N`` skips the prompt; ``--clear`` removes the choice).
"""
import argparse
import sys


def cuda(args):
    """Prints information about the CUDA devices (cf. ``myokit opencl``)."""
    from . import CUDA
    try:
        print(CUDA.info(formatted=True))
    except Exception as e:      # library missing / driver error
        print('CUDA back-end unavailable: ' + str(e))
        return 1
    return 0


def cuda_select(args):
    """Stores the preferred device (cf. ``myokit opencl-select``)."""
    from . import CUDA
    w = 70
    print('=' * w)
    print('Myokit CUDA device selection')
    print('=' * w)
    if args.clear:
        CUDA.save_selection(None)
        print('Selection cleared: device 0 will be used.')
        return 0
    try:
        devices = CUDA.available()
    except Exception as e:
        print('CUDA back-end unavailable: ' + str(e))
        return 1
    print('Selected device: ' + str(CUDA.load_selection()))
    print('=' * w)
    if not devices:
        print('No CUDA devices found.')
        return 1
    for d in devices:
        print(d.format())
        print('-' * w)
    choice = args.device
    if choice is None:
        try:
            text = input('Select device [0-%d], or leave empty to keep the'
                         ' current selection: ' % (len(devices) - 1)).strip()
        except EOFError:
            text = ''
        if text == '':
            print('Selection unchanged.')
            return 0
        try:
            choice = int(text)
        except ValueError:
            print('Invalid selection.')
            return 1
    if choice < 0 or choice >= len(devices):
        print('Invalid selection: there is no device ' + str(choice) + '.')
        return 1
    CUDA.save_selection(choice)
    print('Selected device ' + str(choice) + ': ' + devices[choice].name)
    return 0


def main(argv=None):
    parser = argparse.ArgumentParser(
        prog='python -m myokit_b200',
        description='B200 back-end for Myokit tissue simulations.')
    sub = parser.add_subparsers(dest='command')
    p = sub.add_parser('cuda', help='Prints information about CUDA devices.')
    p.set_defaults(func=cuda)
    p = sub.add_parser('cuda-select', help='Selects the CUDA device to use.')
    p.add_argument('--device', type=int, default=None)
    p.add_argument('--clear', action='store_true')
    p.set_defaults(func=cuda_select)
    args = parser.parse_args(argv)
    if not getattr(args, 'func', None):
        parser.print_help()
        return 2
    return args.func(args)


if __name__ == '__main__':
    sys.exit(main())

"""Updater descriptions -- host mirror of the reference's neural_networks/update_manager.py:3-82.

The reference objects are callables that build lasagne update expressions; here they only carry
the hyper-parameters (and the same ``name`` strings, which are part of the model filename,
update_manager.py:30,42,54,66,79).  The arithmetic runs in the fused CUDA optimizer
(csrc/optim.cu) selected through ``engine_kwargs()``."""


def update_manager_command_parser(parser):
    parser.add_argument('--u_m', dest='update_manager', choices=['adagrad', 'adadelta', 'rmsprop', 'nesterov', 'adam'],
                        help='Update mechanism', default='adam')
    parser.add_argument('--u_l', help='Learning rate', default=0.001, type=float)
    parser.add_argument('--u_rho', help='rho parameter for Adadelta and RMSProp (momentum for Nesterov momentum)',
                        default=0.9, type=float)
    parser.add_argument('--u_b1', help='Beta 1 parameter for Adam', default=0.9, type=float)
    parser.add_argument('--u_b2', help='Beta 2 parameter for Adam', default=0.999, type=float)


def get_update_manager(args):
    kind = args.update_manager
    if kind == 'adagrad':
        return Adagrad(learning_rate=args.u_l)
    if kind == 'adadelta':
        return Adadelta(learning_rate=args.u_l, rho=args.u_rho)
    if kind == 'rmsprop':
        return RMSProp(learning_rate=args.u_l, rho=args.u_rho)
    if kind == 'nesterov':
        return NesterovMomentum(learning_rate=args.u_l, momentum=args.u_rho)
    if kind == 'adam':
        return Adam(learning_rate=args.u_l, beta1=args.u_b1, beta2=args.u_b2)
    raise ValueError('Unknown update option')


class _Updater(object):
    kind = None

    def engine_kwargs(self):
        return dict(updater=self.kind, lr=self.learning_rate, rho=getattr(self, 'rho', getattr(self, 'momentum', 0.9)),
                    beta1=getattr(self, 'beta1', 0.9), beta2=getattr(self, 'beta2', 0.999))


class Adagrad(_Updater):
    kind = 'adagrad'

    def __init__(self, learning_rate=0.1, **kwargs):
        self.learning_rate = learning_rate
        self.name = 'Ug_lr' + str(self.learning_rate)


class Adadelta(_Updater):
    kind = 'adadelta'

    def __init__(self, learning_rate=1.0, rho=0.9, **kwargs):
        self.learning_rate, self.rho = learning_rate, rho
        self.name = 'Ud_lr' + str(self.learning_rate) + '_rho' + str(self.rho)


class RMSProp(_Updater):
    kind = 'rmsprop'

    def __init__(self, learning_rate=1.0, rho=0.9, **kwargs):
        self.learning_rate, self.rho = learning_rate, rho
        self.name = 'Ur_lr' + str(self.learning_rate) + '_rho' + str(self.rho)


class NesterovMomentum(_Updater):
    kind = 'nesterov'

    def __init__(self, learning_rate=1.0, momentum=0.9, **kwargs):
        self.learning_rate, self.momentum = learning_rate, momentum
        self.name = 'Un_lr' + str(self.learning_rate) + '_m' + str(self.momentum)


class Adam(_Updater):
    kind = 'adam'

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, **kwargs):
        self.learning_rate, self.beta1, self.beta2 = learning_rate, beta1, beta2
        self.name = 'Ua_lr' + str(self.learning_rate) + '_b1' + str(self.beta1) + '_b2' + str(self.beta2)

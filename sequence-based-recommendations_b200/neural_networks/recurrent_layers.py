"""Recurrent stack description -- host mirror of neural_networks/recurrent_layers.py:8-39.

The reference's ``RecurrentLayers.__call__`` (recurrent_layers.py:42-104) wires Lasagne layers; in
this build the stack is a handful of integers handed to ``sbr_create``: the cell type, the layer
sizes and the optional embedding.  ``grad_clip`` is always 100 exactly like the reference, whose
``-g`` flag is parsed but never forwarded (command_parser.py:40, recurrent_layers.py:14-15)."""


def recurrent_layers_command_parser(parser):
    parser.add_argument('--r_t', dest='recurrent_layer_type', choices=['LSTM', 'GRU', 'Vanilla'],
                        help='Type of recurrent layer', default='GRU')
    parser.add_argument('--r_l', help="Layers' size, (eg: 100-50-50)", default="50", type=str)
    parser.add_argument('--r_bi', help='Bidirectional layers.', action='store_true')
    parser.add_argument('--r_emb', help='Add an embedding layer before the RNN. Takes the size of the embedding as '
                        'parameter, a size<1 means no embedding layer.', type=int, default=0)


def get_recurrent_layers(args):
    return RecurrentLayers(layer_type=args.recurrent_layer_type, layers=[int(h) for h in args.r_l.split('-')],
                           bidirectional=args.r_bi, embedding_size=args.r_emb)


class RecurrentLayers(object):
    def __init__(self, layer_type="LSTM", layers=(32,), bidirectional=False, embedding_size=0, grad_clipping=100):
        self.layer_type = layer_type
        self.layers = list(layers)
        self.bidirectional = bidirectional
        self.embedding_size = embedding_size
        self.grad_clip = grad_clipping
        self.set_name()

    def set_name(self):
        # part of the checkpoint filename (recurrent_layers.py:28-39)
        name = ""
        if self.bidirectional:
            name += "b" + self.layer_type + "_"
        elif self.layer_type != "LSTM":
            name += self.layer_type + "_"
        name += "gc" + str(self.grad_clip) + "_"
        if self.embedding_size > 0:
            name += "e" + str(self.embedding_size)
        name += "h" + '-'.join(map(str, self.layers))
        self.name = name

    def engine_kwargs(self):
        return dict(cell=self.layer_type, layers=tuple(self.layers), embedding=max(0, self.embedding_size),
                    grad_clip=float(self.grad_clip), bidirectional=bool(self.bidirectional))

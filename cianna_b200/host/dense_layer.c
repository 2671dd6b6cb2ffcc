/*
 * dense_layer.c - dense (fully connected) layer object of the host library.
 *
 * Upstream behaviour followed: src/dense_layer.c:69-326 (creation: "in_size" counts the bias node, weights are
 * W[in_size][nb_neurons + 1] with a pivot column that regenerates the bias node for the next dense layer,
 * silent nb_neurons-1 when nb_neurons % 8 == 0 in mixed precision unless strict_size), :328-438 (save / load),
 * operators of src/cuda/cuda_dense_layer.cu:298-497.
 *
 * Mechanism here: the layer is executed by the convolution kernels with a filter covering the whole input map
 * (see include/cianna_b200.h, "dense"); the flatten order c*A + a of upstream equals the conv column order,
 * so no flatten / reroll buffers exist.  The bias node is the conv "bias column": its constant input value is
 *   - this layer's bias_value when the input comes from the dataset or from a conv / pool / norm layer
 *     (upstream: written by flat_dense, cuda_dense_layer.cu:45-52),
 *   - pivot_weight(previous) * bias_in(previous) when the previous layer is a LINEAR dense layer,
 *   - 0 when the previous layer is a dense layer with a RELU / LOGI / SMAX activation, because those
 *     activations zero the bias node (cuda_activ_functions.cu:48-58) - reproduced on purpose.
 */
#include <math.h>
#include <string.h>
#include "cianna.h"

static const void *dense_input(layer *current)
{
	return current->previous ? current->previous->output : current->c_network->input;
}

static void forward_dense_layer(layer *current)
{
	network *net = current->c_network;
	dense_param *p = (dense_param *)current->param;
	if (net->length == 0) return;
	p->desc.length = net->length;
	CB_CHECK(cb200_conv_forward(&p->desc, &p->w, dense_input(current), current->output, NULL));
	cb_dropout_forward(current);
	if (current->activation_type == SOFTMAX)
		CB_CHECK(cb200_softmax(current->output, net->dtype, net->batch_size, net->length, current->out_c, 1, 1, NULL));
}

static void backward_dense_layer(layer *current)
{
	network *net = current->c_network;
	dense_param *p = (dense_param *)current->param;
	p->desc.length = net->length;
	cb_dropout_backward(current);
	if (current->previous != NULL)
		CB_CHECK(cb200_conv_backward_data(&p->desc, &p->w, current->delta_o, current->previous->delta_o,
			&current->previous->activ, current->previous->output, NULL));
	if (!current->frozen) {
		CB_CHECK(cb200_conv_backward_weights(&p->desc, &p->w, dense_input(current), current->delta_o, NULL));
	}
	cb_dp_layer_done(net, current);
}

/* value carried by the bias node this layer reads */
static float dense_bias_input(layer *current, const float *prev_master)
{
	layer *prev = current->previous;
	if (prev == NULL || prev->type != DENSE) return current->bias_value;
	if (prev->activation_type != LINEAR) return 0.0f;
	{
		dense_param *pp = (dense_param *)prev->param;
		float pivot = prev_master[(size_t)pp->in_size * (pp->nb_neurons + 1) - 1];
		return pivot * pp->desc.bias_value;
	}
}

int dense_create(network *net, layer *previous, int nb_neurons, const char *activation, float *bias,
	float drop_rate, int strict_size, const char *init_fct, float init_scaling, FILE *f_load, int f_bin)
{
	layer *current;
	dense_param *p;
	float *host_w, *prev_master = NULL;
	size_t nw, i;
	int j, pc, ph, pw;
	char activ[40];

	if (f_load == NULL && !strict_size && net->use_cuda_TC != FP32C_FP32A && nb_neurons % 8 == 0)
		nb_neurons -= 1;
	if (net->nb_layers >= MAX_LAYERS_NB) { printf("\nERROR: too many layers\n"); exit(EXIT_FAILURE); }
	current = (layer *)calloc(1, sizeof(layer));
	current->index = net->nb_layers;
	net->net_layers[net->nb_layers++] = current;
	current->c_network = net;
	current->type = DENSE;
	current->previous = previous;
	printf("L:%d - CREATING DENSE LAYER ...\n", net->nb_layers);
	current->dropout_rate = drop_rate;
	load_activ_param(current, activation);

	p = (dense_param *)calloc(1, sizeof(dense_param));
	p->nb_neurons = nb_neurons;
	if (previous == NULL) { pc = net->in_dims[3]; ph = net->in_dims[1] * net->in_dims[2]; pw = net->in_dims[0]; }
	else { pc = previous->out_c; ph = previous->out_h; pw = previous->out_w; }
	p->prev_c = pc; p->prev_h = ph; p->prev_w = pw;
	p->in_size = pc * ph * pw + 1;
	current->out_c = nb_neurons; current->out_h = 1; current->out_w = 1; current->out_d = 1;
	current->param = p;
	set_activ_defaults(current, activation);
	if (bias != NULL) current->bias_value = *bias;
	if (previous == NULL) current->bias_value = net->input_bias;
	if (current->activation_type == YOLO) { printf("\nERROR: YOLO activation on a dense layer is not supported.\n"); exit(EXIT_FAILURE); }

	nw = (size_t)p->in_size * (nb_neurons + 1);
	host_w = (float *)calloc(nw, sizeof(float));
	if (f_load == NULL) {
		/* the layer that follows a dense layer sets the previous pivot weight (src/dense_layer.c:209-228) */
		if (previous != NULL && previous->type == DENSE) {
			dense_param *pp = (dense_param *)previous->param;
			float pivot = current->bias_value / previous->bias_value;
			CB_CHECK(cb200_h2d(pp->w.master + ((size_t)pp->in_size * (pp->nb_neurons + 1) - 1), &pivot, sizeof(float), NULL));
			CB_CHECK(cb200_stream_sync(NULL));
		}
		{
			/* fan-in rows x neurons, pivot column left at zero */
			float *tmp = (float *)calloc((size_t)p->in_size * nb_neurons, sizeof(float));
			init_weights(tmp, p->in_size, nb_neurons, init_fct, init_scaling);
			for (i = 0; i < (size_t)p->in_size; i++)
				for (j = 0; j < nb_neurons; j++)
					host_w[i * (nb_neurons + 1) + j] = tmp[i * nb_neurons + j];
			free(tmp);
		}
	} else if (f_bin) {
		fread(host_w, sizeof(float), nw, f_load);
	} else {
		for (i = 0; i < nw; i++) fscanf(f_load, "%f", &host_w[i]);
	}
	if (previous != NULL && previous->type == DENSE) {
		dense_param *pp = (dense_param *)previous->param;
		size_t pn = (size_t)pp->in_size * (pp->nb_neurons + 1);
		prev_master = (float *)malloc(pn * sizeof(float));
		CB_CHECK(cb200_d2h(prev_master, pp->w.master, pn * sizeof(float), NULL));
		CB_CHECK(cb200_stream_sync(NULL));
	}

	p->desc.dtype = net->dtype; p->desc.batch = net->batch_size; p->desc.length = net->batch_size;
	p->desc.in_c = pc; p->desc.in_h = ph; p->desc.in_w = pw;
	p->desc.out_c = nb_neurons; p->desc.out_h = 1; p->desc.out_w = 1;
	p->desc.f_h = ph; p->desc.f_w = pw; p->desc.stride_h = 1; p->desc.stride_w = 1; p->desc.pad_h = 0; p->desc.pad_w = 0;
	p->desc.bias_value = dense_bias_input(current, prev_master);
	p->desc.activ = current->activ;
	if (current->activation_type == SOFTMAX) p->desc.activ.type = CB200_LINEAR;
	if (drop_rate > 0.01f) p->desc.activ.type = CB200_LINEAR;   /* the activation runs in the dropout pass, after the mask */
	cb_dropout_setup(current);
	free(prev_master);

	{
		size_t es = cb200_dtype_size(net->dtype);
		CB_CHECK(cb200_malloc((void **)&p->w.master, nw * sizeof(float)));
		CB_CHECK(cb200_malloc(&p->w.w_fwd, cb200_conv_wfwd_elems(&p->desc) * es));
		CB_CHECK(cb200_malloc(&p->w.w_bwd, cb200_conv_wbwd_elems(&p->desc) * es));
		CB_CHECK(cb200_malloc((void **)&p->w.bias_w, (size_t)nb_neurons * sizeof(float)));
		if (!net->inference_only) CB_CHECK(cb200_malloc((void **)&p->w.moment, nw * sizeof(float)));
		CB_CHECK(cb200_malloc(&current->output, (size_t)net->batch_size * cb200_round_channels(nb_neurons) * es));
		if (!net->inference_only) CB_CHECK(cb200_malloc(&current->delta_o, (size_t)net->batch_size * cb200_round_channels(nb_neurons) * es));
	}
	CB_CHECK(cb200_h2d(p->w.master, host_w, nw * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
	free(host_w);
	CB_CHECK(cb200_dense_prepare_weights(&p->desc, &p->w, NULL));

	current->forward = forward_dense_layer;
	current->backprop = backward_dense_layer;
	current->nb_params = p->in_size * (nb_neurons + 1);
	print_string_activ_param(current, activ);
	printf("      Input: %d, Nb. Neurons: %d\n      Activation: %s, Bias: %0.2f, dropout rate: %0.2f\n      Nb. weights: %d\n",
		p->in_size, p->nb_neurons, activ, current->bias_value, current->dropout_rate, (p->nb_neurons + 1) * p->in_size);
	net->total_nb_param += (long long)(p->nb_neurons + 1) * p->in_size;
	return net->nb_layers - 1;
}

void dense_get_weights(layer *cur, float *dst, int moment)
{
	dense_param *p = (dense_param *)cur->param;
	CB_CHECK(cb200_d2h(dst, moment ? p->w.moment : p->w.master, (size_t)p->in_size * (p->nb_neurons + 1) * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
}

void dense_set_weights(layer *cur, const float *src)
{
	dense_param *p = (dense_param *)cur->param;
	network *net = cur->c_network;
	int k;
	CB_CHECK(cb200_h2d(p->w.master, src, (size_t)p->in_size * (p->nb_neurons + 1) * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
	CB_CHECK(cb200_dense_prepare_weights(&p->desc, &p->w, NULL));
	/* a dense layer reading this one's bias node holds pivot * bias_in as a constant (dense_bias_input): the pivot may
	 * just have changed.  (Limitation kept: behind a LINEAR dense layer upstream's bias node is a trained output column -
	 * its non-pivot rows receive gradient - whereas here it stays the constant pivot * bias_in.) */
	for (k = cur->index + 1; k < net->nb_layers; k++) {
		layer *next = net->net_layers[k];
		if (next->type == DENSE && next->previous == cur) {
			dense_param *np = (dense_param *)next->param;
			np->desc.bias_value = dense_bias_input(next, src);
		}
	}
}

/* rebuild the 16-bit operand copies after the FP32 master changed in place (data-parallel parameter broadcast) */
void dense_refresh_operands(layer *cur)
{
	dense_param *p = (dense_param *)cur->param;
	CB_CHECK(cb200_dense_prepare_weights(&p->desc, &p->w, NULL));
}

void dense_save(FILE *f, layer *current, int f_bin)
{
	dense_param *p = (dense_param *)current->param;
	char layer_type = 'D';
	size_t nw = (size_t)p->in_size * (p->nb_neurons + 1), i;
	float *host_w = (float *)malloc(nw * sizeof(float));
	int j;
	if (f_bin) {
		fwrite(&layer_type, sizeof(char), 1, f);
		fwrite(&p->nb_neurons, sizeof(int), 1, f);
		fwrite(&current->dropout_rate, sizeof(float), 1, f);
		fwrite(&current->bias_value, sizeof(float), 1, f);
		print_activ_param(f, current, f_bin);
	} else {
		fprintf(f, "D");
		fprintf(f, "%dn%fd%fb", p->nb_neurons, current->dropout_rate, current->bias_value);
		print_activ_param(f, current, f_bin);
		fprintf(f, "\n");
	}
	dense_get_weights(current, host_w, 0);
	if (f_bin) {
		fwrite(host_w, sizeof(float), nw, f);
	} else {
		for (i = 0; i < (size_t)p->in_size; i++) {
			for (j = 0; j < p->nb_neurons + 1; j++) fprintf(f, "%g ", host_w[i * (p->nb_neurons + 1) + j]);
			fprintf(f, "\n");
		}
		fprintf(f, "\n");
	}
	free(host_w);
}

void dense_load(network *net, FILE *f, int f_bin)
{
	int nb_neurons;
	float dropout_rate, bias;
	char activ_type[40];
	layer *previous;
	printf("Loading dense layer, L:%d\n", net->nb_layers + 1);
	if (f_bin) {
		fread(&nb_neurons, sizeof(int), 1, f);
		fread(&dropout_rate, sizeof(float), 1, f);
		fread(&bias, sizeof(float), 1, f);
		fread(activ_type, sizeof(char), 40, f);
	} else {
		fscanf(f, "%dn%fd%fb%s\n", &nb_neurons, &dropout_rate, &bias, activ_type);
	}
	previous = net->nb_layers <= 0 ? NULL : net->net_layers[net->nb_layers - 1];
	dense_create(net, previous, nb_neurons, activ_type, &bias, dropout_rate, 1, NULL, 0.0f, f, f_bin);
}

/*
 * layers.c - conv / pool / norm layer objects of the host library: shape inference, device
 * allocation, weight init/load, save, and the forward/backprop operators that drive the C-ABI.
 *
 * Behavioural contract followed (upstream): src/conv_layer.c:80-417 (conv_create), :420-677 (save/load),
 * src/pool_layer.c:111-287, src/norm_layer.c:110-375; operator contract of SURVEY.md 8b:
 *   forward(l)  reads l->previous->output (or net->input), writes l->output, activation included;
 *   backprop(l) consumes l->delta_o, writes l->previous->delta_o already multiplied by the previous
 *               layer's activation derivative, then produces this layer's raw gradients
 *               (the optimizer runs after the backward sweep so gradients can be all-reduced first).
 */
#include <math.h>
#include <string.h>
#include "cianna.h"

/* ------------------------------------------------------------------ helpers */
static size_t act_bytes(network *net, int c, int h, int w)
{
	return (size_t)net->batch_size * h * w * cb200_round_channels(c) * cb200_dtype_size(net->dtype);
}

static void *dev_alloc(size_t bytes)
{
	void *p = NULL;
	CB_CHECK(cb200_malloc(&p, bytes));
	return p;
}

/* shape (c,h,w) of what a layer hands to its successor */
/* channels, rows (depth x height) and width of the map a layer reads; prev_depth() gives the depth alone */
static void prev_shape(network *net, layer *previous, int *c, int *h, int *w)
{
	if (previous == NULL) { *c = net->in_dims[3]; *h = net->in_dims[1] * net->in_dims[2]; *w = net->in_dims[0]; }
	else { *c = previous->out_c; *h = previous->out_h; *w = previous->out_w; }
}

static int prev_depth(network *net, layer *previous)
{
	int d = previous == NULL ? net->in_dims[2] : previous->out_d;
	return d > 0 ? d : 1;
}

static const void *layer_input(layer *current)
{
	return current->previous ? current->previous->output : current->c_network->input;
}

int nb_area_comp(int size, int f_size, int padding, int int_padding, int stride)
{
	if ((size + padding * 2 - f_size) % stride != 0)
		printf(" WARNING: unable to divide current input volume into an integer number of conv/pool regions\n This might produce unstable results !\n\n");
	return (size + (size - 1) * int_padding + padding * 2 - f_size) / stride + 1;
}

/* ---- group-norm + max-pool fusion ---- */
static int fusion_enabled = -1;
void cb_set_fusion(int on) { fusion_enabled = on ? 1 : 0; }
static int fusion_on(void)
{
	if (fusion_enabled < 0) {
		const char *e = getenv("CB200_NO_FUSION");
		fusion_enabled = (e != NULL && e[0] != '\0' && e[0] != '0') ? 0 : 1;
	}
	return fusion_enabled;
}

/* a fused norm layer keeps no full-resolution output / delta; anything else that wants to read them un-fuses it */
static void unfuse_norm(layer *norm)
{
	network *net = norm->c_network;
	norm_param *np = (norm_param *)norm->param;
	if (np->fused_pool == NULL) return;
	((pool_param *)np->fused_pool->param)->fused_norm = 0;
	np->fused_pool = NULL;
	norm->output = dev_alloc(act_bytes(net, norm->out_c, norm->out_h, norm->out_w));
	if (!net->inference_only) norm->delta_o = dev_alloc(act_bytes(net, norm->out_c, norm->out_h, norm->out_w));
}

static layer *new_layer(network *net, int type, layer *previous)
{
	layer *current = (layer *)calloc(1, sizeof(layer));
	if (previous != NULL && previous->type == NORM) unfuse_norm(previous);
	if (net->nb_layers >= MAX_LAYERS_NB) { printf("\nERROR: too many layers (MAX_LAYERS_NB=%d)\n", MAX_LAYERS_NB); exit(EXIT_FAILURE); }
	current->index = net->nb_layers;
	net->net_layers[net->nb_layers++] = current;
	current->c_network = net;
	current->type = type;
	current->previous = previous;
	current->out_d = previous == NULL ? (net->in_dims[2] > 0 ? net->in_dims[2] : 1) : previous->out_d;   /* conv / pool / dense set their own */
	return current;
}

/* ------------------------------------------------------------------ dropout (shared by conv / pool / dense layers)
 * Call order of upstream kept (src/cuda/cuda_conv_layer.cu:399-421,440-447 and the dense / pool twins): mask or scale
 * the pre-activation output, then activate; mask the delta before anything else in the backward pass. */
static int drop_on(layer *current) { return current->dropout_rate > 0.01f; }
static int drop_draws_mask(network *net) { return net->is_inference == 0 || net->inference_drop_mode == MC_MODEL; }

void cb_dropout_setup(layer *current)
{
	network *net = current->c_network;
	cb200_dropout_desc *d = &current->drop;
	if (!drop_on(current)) return;
	d->dtype = net->dtype; d->batch = net->batch_size; d->length = net->batch_size;
	d->c = current->out_c; d->h = current->out_h; d->w = current->out_w;
	d->drop_rate = current->dropout_rate;
	d->stream_id = (unsigned int)current->index;
	d->activ = current->activ;
	if (current->activation_type == SOFTMAX || current->activation_type == YOLO) d->activ.type = CB200_LINEAR;
}

void cb_dropout_forward(layer *current)
{
	network *net = current->c_network;
	cb200_dropout_desc *d = &current->drop;
	if (!drop_on(current)) return;
	d->length = net->length;
	if (drop_draws_mask(net)) {
		d->seed = net->drop_seed;
		d->draw = ++net->drop_draw;
		CB_CHECK(cb200_dropout_forward(d, current->output, 0, NULL));
	} else
		CB_CHECK(cb200_dropout_forward(d, current->output, 1, NULL));
}

void cb_dropout_backward(layer *current)
{
	if (!drop_on(current) || !drop_draws_mask(current->c_network)) return;
	CB_CHECK(cb200_dropout_backward(&current->drop, current->delta_o, NULL));
}

void cb_set_dropout_seed(network *net, unsigned long long seed) { net->drop_seed = seed; }
/* AVG_MODEL (0) / MC_MODEL (1) for cb_forward(.., is_inference = 1); forward_testset takes it as an argument like upstream */
void cb_set_inference_drop_mode(network *net, int mode) { net->inference_drop_mode = mode ? MC_MODEL : AVG_MODEL; }

/* ------------------------------------------------------------------ convolution */
static void forward_conv_layer(layer *current)
{
	network *net = current->c_network;
	conv_param *p = (conv_param *)current->param;
	if (net->length == 0) return;
	p->desc.length = net->length;
	{
		/* a group-norm layer next: its statistics come out of this layer's epilogue when the kernel can (no dropout in
		 * between - it would change the tensor the statistics are about) */
		layer *next = current->index + 1 < net->nb_layers ? net->net_layers[current->index + 1] : NULL;
		if (next != NULL && next->type == NORM && next->previous == current && !drop_on(current)
			&& current->activation_type != SOFTMAX && current->activation_type != YOLO) {
			norm_param *np = (norm_param *)next->param;
			int done = 0;
			np->desc.length = net->length;
			CB_CHECK(cb200_conv_forward_stats(&p->desc, &p->w, layer_input(current), current->output, &np->desc, np->workspace, &done, NULL));
			np->stats_ready = done;
		} else
			CB_CHECK(cb200_conv_forward(&p->desc, &p->w, layer_input(current), current->output, NULL));
	}
	cb_dropout_forward(current);
	if (current->activation_type == SOFTMAX)
		CB_CHECK(cb200_softmax(current->output, net->dtype, net->batch_size, net->length, current->out_c, current->out_h, current->out_w, NULL));
	else if (current->activation_type == YOLO) {
		yolo_param *y = (yolo_param *)current->activ_param;
		y->desc.length = net->length;
		CB_CHECK(cb200_yolo_activation(&y->desc, current->output, NULL));
	}
}

static void backward_conv_layer(layer *current)
{
	network *net = current->c_network;
	conv_param *p = (conv_param *)current->param;
	p->desc.length = net->length;
	cb_dropout_backward(current);
	/* the weight gradient only needs what is already enqueued (this layer's delta and, when the following norm layer
	 * produced it, grad_b): it goes to the low-priority side stream, ordered after that point, BEFORE the data gradient
	 * is enqueued on the compute stream, so the critical path never waits for it; joined again before the optimizer
	 * (network.c: apply_updates) */
	if (!current->frozen && net->wgrad_stream != NULL) CB_CHECK(cb200_stream_wait(net->wgrad_stream, NULL));
	if (current->previous != NULL)
		CB_CHECK(cb200_conv_backward_data(&p->desc, &p->w, current->delta_o, current->previous->delta_o,
			&current->previous->activ, current->previous->output, NULL));
	if (!current->frozen) {
		void *ws = net->wgrad_stream;
		CB_CHECK(cb200_conv_backward_weights_ex(&p->desc, &p->w, layer_input(current), current->delta_o, p->bias_grad_from_next, ws));
	}
	cb_dp_layer_done(net, current);
}

static void conv_alloc_weights(network *net, cb200_conv_desc *d, cb200_conv_weights *w)
{
	size_t es = cb200_dtype_size(net->dtype);
	w->master = (float *)dev_alloc(cb200_conv_master_elems(d) * sizeof(float));
	w->w_fwd = dev_alloc(cb200_conv_wfwd_elems(d) * es);
	w->bias_w = (float *)dev_alloc((size_t)d->out_c * sizeof(float));
	if (!net->inference_only) {
		w->moment = (float *)dev_alloc(cb200_conv_master_elems(d) * sizeof(float));
		w->w_bwd = dev_alloc(cb200_conv_wbwd_elems(d) * es);
	} else {
		/* the operand builder writes both layouts; give it a scratch target */
		w->moment = NULL;
		w->w_bwd = dev_alloc(cb200_conv_wbwd_elems(d) * es);
	}
	w->grad = NULL;
	w->grad_b = NULL;
}

int conv_create(network *net, layer *previous, int *f_size, int nb_filters, int *stride, int *padding,
	int *int_padding, int *in_shape, const char *activation, float *bias, float drop_rate,
	const char *init_fct, float init_scaling, FILE *f_load, int f_bin)
{
	int k, pc, ph, pw, pd, generic;
	layer *current = new_layer(net, CONV, previous);
	conv_param *p = (conv_param *)calloc(1, sizeof(conv_param));
	float *host_w;
	size_t nw;
	char activ[40];

	printf("L:%d - CREATING CONVOLUTIONAL LAYER ...\n", net->nb_layers);
	load_activ_param(current, activation);
	for (k = 0; k < 3; k++) {
		if (stride[k] > f_size[k]) { printf("\nERROR: filter size cannot be smaller than stride size in a given dimension !\n"); exit(EXIT_FAILURE); }
		p->f_size[k] = f_size[k]; p->stride[k] = stride[k]; p->padding[k] = padding[k]; p->int_padding[k] = int_padding[k];
	}
	p->nb_filters = nb_filters;
	current->dropout_rate = drop_rate;

	if (previous != NULL && previous->type == DENSE) {
		if (in_shape == NULL) { printf("ERROR: dense to conv conversion requires input_shape to be defined in conv_layer.\n\n"); exit(EXIT_FAILURE); }
		printf("\nERROR: dense to conv stacking is not supported by the B200 core yet.\n"); exit(EXIT_FAILURE);
	}
	prev_shape(net, previous, &pc, &ph, &pw);
	pd = prev_depth(net, previous);
	ph /= pd;      /* true height */
	p->prev_size[0] = pw; p->prev_size[1] = ph; p->prev_size[2] = pd; p->prev_depth = pc;
	p->flat_f_size = f_size[0] * f_size[1] * f_size[2] * pc + 1;
	for (k = 0; k < 3; k++)
		p->nb_area[k] = nb_area_comp(p->prev_size[k], p->f_size[k], p->padding[k], p->int_padding[k], p->stride[k]);
	/* three-dimensional maps and internal padding (transposed convolution) run on the generic CUDA-core kernels */
	generic = pd > 1 || p->nb_area[2] > 1 || f_size[2] > 1 || padding[2] > 0 || int_padding[0] > 0 || int_padding[1] > 0 || int_padding[2] > 0;

	current->out_c = nb_filters; current->out_w = p->nb_area[0]; current->out_d = p->nb_area[2];
	current->out_h = p->nb_area[1] * p->nb_area[2];
	current->param = p;
	set_activ_defaults(current, activation);
	if (bias != NULL) current->bias_value = *bias;
	if (previous == NULL) current->bias_value = net->input_bias;
	if (current->activation_type == YOLO) set_yolo_activ(current);   /* checks nb_filters against the YOLO set-up */

	p->desc.dtype = net->dtype; p->desc.batch = net->batch_size; p->desc.length = net->batch_size;
	p->desc.in_c = pc; p->desc.in_h = ph; p->desc.in_w = pw;
	p->desc.out_c = nb_filters; p->desc.out_h = p->nb_area[1]; p->desc.out_w = current->out_w;
	p->desc.f_h = f_size[1]; p->desc.f_w = f_size[0];
	p->desc.stride_h = stride[1]; p->desc.stride_w = stride[0];
	p->desc.pad_h = padding[1]; p->desc.pad_w = padding[0];
	if (generic) {
		p->desc.in_d = pd; p->desc.out_d = p->nb_area[2]; p->desc.f_d = f_size[2]; p->desc.stride_d = stride[2]; p->desc.pad_d = padding[2];
		p->desc.ipad_w = int_padding[0]; p->desc.ipad_h = int_padding[1]; p->desc.ipad_d = int_padding[2];
	}
	p->desc.bias_value = current->bias_value;
	p->desc.activ = current->activ;
	if (current->activation_type == SOFTMAX || current->activation_type == YOLO)
		p->desc.activ.type = CB200_LINEAR;   /* softmax / the YOLO head are separate passes on the linear output */
	if (drop_rate > 0.01f) p->desc.activ.type = CB200_LINEAR;   /* the activation runs in the dropout pass, after the mask */
	cb_dropout_setup(current);
	/* first layer on an input with very few channels (RGB / grey): the layout import unrolls the receptive fields into
	 * patch rows so that the layer runs on the tensor-core GEMM kernels (include/cianna_b200.h, cb200_import_input_patches) */
	p->desc.input_is_patches = (previous == NULL && cb200_round_channels(pc) < 16 && !generic) ? 1 : 0;
	if (p->desc.input_is_patches && cb200_conv_first_direct(&p->desc)) {
		/* ... or, for the usual first-layer shapes, the kernels build those rows in shared memory straight from the
		 * dataset batch: no import pass, no patch tensor in HBM; the layer's input pointer is the batch itself */
		p->desc.input_is_patches = 2;
		CB_CHECK(cb200_free(net->input));
		net->input = NULL;
		net->patch_desc = &p->desc;
	} else if (p->desc.input_is_patches) {
		size_t bytes = (size_t)net->batch_size * current->out_h * current->out_w * cb200_patch_width(pc, f_size[1], f_size[0]) * cb200_dtype_size(net->dtype);
		CB_CHECK(cb200_free(net->input));
		net->input = dev_alloc(bytes);
		net->patch_desc = &p->desc;
	}

	conv_alloc_weights(net, &p->desc, &p->w);
	current->output = dev_alloc(act_bytes(net, current->out_c, current->out_h, current->out_w));
	if (!net->inference_only)
		current->delta_o = dev_alloc(act_bytes(net, current->out_c, current->out_h, current->out_w));

	nw = (size_t)nb_filters * p->flat_f_size;
	host_w = (float *)calloc(nw, sizeof(float));
	if (f_load == NULL) {
		init_weights(host_w, p->flat_f_size, nb_filters, init_fct, init_scaling);
	} else if (f_bin) {
		fread(host_w, sizeof(float), nw, f_load);
	} else {
		size_t i;
		for (i = 0; i < nw; i++) fscanf(f_load, "%f", &host_w[i]);
	}
	CB_CHECK(cb200_h2d(p->w.master, host_w, nw * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
	free(host_w);
	CB_CHECK(cb200_conv_prepare_weights(&p->desc, &p->w, NULL));

	current->forward = forward_conv_layer;
	current->backprop = backward_conv_layer;
	current->nb_params = nb_filters * p->flat_f_size;

	print_string_activ_param(current, activ);
	printf("      Input: %dx%dx%dx%d, Filters: %df %dx%dx%dx%d, Output: %dx%dx%dx%d \n"
	       "      Stride: %d:%d:%d, padding: %d:%d:%d, int_padding: %d:%d:%d,  \n"
	       "      Activation: %s, Bias: %0.2f, dropout rate: %0.2f\n"
	       "      Nb. weights: %d\n",
		p->prev_size[0], p->prev_size[1], p->prev_size[2], p->prev_depth, p->nb_filters,
		p->f_size[0], p->f_size[1], p->f_size[2], p->prev_depth,
		p->nb_area[0], p->nb_area[1], p->nb_area[2], p->nb_filters,
		p->stride[0], p->stride[1], p->stride[2], p->padding[0], p->padding[1], p->padding[2],
		p->int_padding[0], p->int_padding[1], p->int_padding[2],
		activ, current->bias_value, current->dropout_rate, nb_filters * p->flat_f_size);
	net->total_nb_param += nb_filters * p->flat_f_size;
	return net->nb_layers - 1;
}

void conv_save(FILE *f, layer *current, int f_bin)
{
	conv_param *p = (conv_param *)current->param;
	char layer_type = 'C';
	size_t nw = (size_t)p->nb_filters * p->flat_f_size, i;
	float *host_w = (float *)malloc(nw * sizeof(float));
	int j;

	if (f_bin) {
		fwrite(&layer_type, sizeof(char), 1, f);
		fwrite(&p->nb_filters, sizeof(int), 1, f);
		fwrite(p->f_size, sizeof(int), 3, f);
		fwrite(p->stride, sizeof(int), 3, f);
		fwrite(p->padding, sizeof(int), 3, f);
		fwrite(p->int_padding, sizeof(int), 3, f);
		fwrite(p->prev_size, sizeof(int), 3, f);
		fwrite(&p->prev_depth, sizeof(int), 1, f);
		fwrite(&current->dropout_rate, sizeof(float), 1, f);
		fwrite(&current->bias_value, sizeof(float), 1, f);
		print_activ_param(f, current, f_bin);
		if (current->activation_type == YOLO) {
			/* YOLO set-up block of the save format (src/conv_layer.c:442-456): priors dimension-major */
			yolo_param *y = current->c_network->y_param;
			int a, b;
			fwrite(&y->nb_box, sizeof(int), 1, f);
			fwrite(&y->nb_class, sizeof(int), 1, f);
			fwrite(&y->nb_param, sizeof(int), 1, f);
			fwrite(&y->fit_dim, sizeof(int), 1, f);
			fwrite(&y->class_softmax, sizeof(int), 1, f);
			for (a = 0; a < 3; a++)
				for (b = 0; b < y->nb_box; b++) fwrite(y->prior_size + b * 3 + a, sizeof(float), 1, f);
			for (a = 0; a < 6; a++) fwrite(y->slopes_and_maxes_tab[a], sizeof(float), 3, f);
		}
	} else {
		fprintf(f, "C");
		fprintf(f, "%df%dx%dx%d.%dx%dx%ds%dx%dx%dp%dx%dx%dip%dx%dx%dx%didim%fd%fb", p->nb_filters,
			p->f_size[0], p->f_size[1], p->f_size[2], p->stride[0], p->stride[1], p->stride[2],
			p->padding[0], p->padding[1], p->padding[2], p->int_padding[0], p->int_padding[1], p->int_padding[2],
			p->prev_size[0], p->prev_size[1], p->prev_size[2], p->prev_depth, current->dropout_rate, current->bias_value);
		print_activ_param(f, current, f_bin);
		fprintf(f, "\n");
		if (current->activation_type == YOLO) {
			yolo_param *y = current->c_network->y_param;
			int a, b;
			fprintf(f, "%d %d %d %d %d\n", y->nb_box, y->nb_class, y->nb_param, y->fit_dim, y->class_softmax);
			for (a = 0; a < 3; a++) {
				for (b = 0; b < y->nb_box; b++) fprintf(f, "%g ", y->prior_size[b * 3 + a]);
				fprintf(f, "\n");
			}
			for (a = 0; a < 6; a++)
				fprintf(f, "%g %g %g \n", y->slopes_and_maxes_tab[a][0], y->slopes_and_maxes_tab[a][1], y->slopes_and_maxes_tab[a][2]);
		}
	}
	CB_CHECK(cb200_d2h(host_w, p->w.master, nw * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
	if (f_bin) {
		fwrite(host_w, sizeof(float), nw, f);
	} else {
		for (i = 0; i < (size_t)p->nb_filters; i++) {
			for (j = 0; j < p->flat_f_size; j++) fprintf(f, "%g ", host_w[i * p->flat_f_size + j]);
			fprintf(f, "\n");
		}
		fprintf(f, "\n");
	}
	free(host_w);
}

void conv_load(network *net, FILE *f, int f_bin)
{
	int nb_filters, f_size[3], stride[3], padding[3], int_padding[3], input_shape[4];
	float dropout_rate, bias;
	char activ_type[40];
	layer *previous;

	printf("Loading conv layer, L:%d\n", net->nb_layers + 1);
	if (f_bin) {
		fread(&nb_filters, sizeof(int), 1, f);
		fread(f_size, sizeof(int), 3, f);
		fread(stride, sizeof(int), 3, f);
		fread(padding, sizeof(int), 3, f);
		fread(int_padding, sizeof(int), 3, f);
		fread(input_shape, sizeof(int), 4, f);
		fread(&dropout_rate, sizeof(float), 1, f);
		fread(&bias, sizeof(float), 1, f);
		fread(activ_type, sizeof(char), 40, f);
	} else {
		fscanf(f, "%df%dx%dx%d.%dx%dx%ds%dx%dx%dp%dx%dx%dip%dx%dx%dx%didim%fd%fb%s\n", &nb_filters,
			&f_size[0], &f_size[1], &f_size[2], &stride[0], &stride[1], &stride[2],
			&padding[0], &padding[1], &padding[2], &int_padding[0], &int_padding[1], &int_padding[2],
			&input_shape[0], &input_shape[1], &input_shape[2], &input_shape[3], &dropout_rate, &bias, activ_type);
	}
	if (strncmp(activ_type, "YOLO", 4) == 0) {
		/* YOLO set-up block (src/conv_layer.c:571-650): the saved geometry of the head replaces the one given to
		 * set_yolo_params unless no_override was requested */
		int head[5], a, b;
		float *prior, sm[6][3];
		if (f_bin) fread(head, sizeof(int), 5, f);
		else fscanf(f, "%d %d %d %d %d\n", &head[0], &head[1], &head[2], &head[3], &head[4]);
		if (head[0] <= 0 || head[0] > CB200_YOLO_MAX_BOX) { printf("\nERROR: corrupted YOLO block in save file (nb_box = %d)\n", head[0]); exit(EXIT_FAILURE); }
		prior = (float *)calloc(3 * head[0], sizeof(float));
		for (a = 0; a < 3; a++)
			for (b = 0; b < head[0]; b++) {
				if (f_bin) fread(prior + b * 3 + a, sizeof(float), 1, f);
				else fscanf(f, "%f", prior + b * 3 + a);
			}
		for (a = 0; a < 6; a++) {
			if (f_bin) fread(sm[a], sizeof(float), 3, f);
			else fscanf(f, "%f %f %f", &sm[a][0], &sm[a][1], &sm[a][2]);
		}
		if (net->y_param->no_override != 1) {
			yolo_param *y = net->y_param;
			y->nb_box = head[0]; y->nb_class = head[1]; y->nb_param = head[2]; y->fit_dim = head[3]; y->class_softmax = head[4];
			free(y->prior_size);
			y->prior_size = prior;
			memcpy(y->slopes_and_maxes_tab, sm, sizeof(sm));
			printf(" WARNING: Overriding the following YOLO parameters from save file:\n");
			printf(" Nboxes = %d, Nclasses = %d, Nparams = %d\n", y->nb_box, y->nb_class, y->nb_param);
			printf(" Classification type: %s\n", y->class_softmax ? "softmax-CrossEntropy" : "sigmoid-MSE");
			printf(" Nb dim fitted : %d\n\n", y->fit_dim);
			for (a = 0; a < 3; a++) {
				printf(" %c priors = [", "WHD"[a]);
				for (b = 0; b < y->nb_box; b++) printf("%4.4f ", y->prior_size[b * 3 + a]);
				printf("]\n");
			}
			printf("\n Activation slopes and limits: \n   = ");
			for (a = 0; a < 6; a++) printf("[%6.2f %6.2f %6.2f]\n     ", sm[a][0], sm[a][1], sm[a][2]);
			printf("\n");
		} else free(prior);
	}
	previous = net->nb_layers <= 0 ? NULL : net->net_layers[net->nb_layers - 1];
	conv_create(net, previous, f_size, nb_filters, stride, padding, int_padding, input_shape, activ_type, &bias, dropout_rate, NULL, 0.0f, f, f_bin);
}

/* ------------------------------------------------------------------ pooling */
static void forward_pool_layer(layer *current)
{
	network *net = current->c_network;
	pool_param *p = (pool_param *)current->param;
	if (net->length == 0) return;
	p->desc.length = net->length;
	if (p->fused_norm) {
		layer *norm = current->previous;
		norm_param *np = (norm_param *)norm->param;
		np->desc.length = net->length;
		CB_CHECK(cb200_norm_pool_forward_ex(&np->desc, &p->desc, norm->previous->output, current->output, p->pool_map,
			np->gamma, np->beta, np->mean, np->var, np->workspace, np->stats_ready, NULL));
		np->stats_ready = 0;
	} else
		CB_CHECK(cb200_pool_forward(&p->desc, layer_input(current), current->output, p->pool_map, NULL));
	cb_dropout_forward(current);
	if (current->activation_type == SOFTMAX)
		CB_CHECK(cb200_softmax(current->output, net->dtype, net->batch_size, net->length, current->out_c, current->out_h, current->out_w, NULL));
}

static void backward_pool_layer(layer *current)
{
	network *net = current->c_network;
	pool_param *p = (pool_param *)current->param;
	p->desc.length = net->length;
	cb_dropout_backward(current);
	if (p->fused_norm) return;      /* the norm layer's backward consumes this layer's delta and map directly */
	if (current->previous != NULL)
		CB_CHECK(cb200_pool_backward(&p->desc, current->delta_o, p->pool_map, current->previous->delta_o,
			&current->previous->activ, current->previous->output, NULL));
}

int pool_create(network *net, layer *previous, int *pool_size, int *stride, int *padding,
	const char *char_pool_type, const char *activation, int global, float drop_rate)
{
	int k, pc, ph, pw, pd;
	layer *current = new_layer(net, POOL, previous);
	pool_param *p = (pool_param *)calloc(1, sizeof(pool_param));
	char activ[40];

	printf("L:%d - CREATING POOL LAYER ...\n", net->nb_layers);
	load_activ_param(current, activation);
	current->dropout_rate = drop_rate;
	if (previous != NULL && previous->dropout_rate > 0.01f) {
		printf("\nERROR: A pooling layer cannot be set if dropout is used in the previous layer due to problem with weight/output rescaling.\n");
		exit(EXIT_FAILURE);
	}
	for (k = 0; k < 3; k++) {
		if (stride[k] > pool_size[k]) { printf("\nERROR: pool size cannot be smaller than stride size in a given dimension !\n"); exit(EXIT_FAILURE); }
		if (padding[k] > pool_size[k]) { printf("\nERROR: pool size cannot be equal or smaller than padding in a given dimension !\n"); exit(EXIT_FAILURE); }
		p->p_size[k] = pool_size[k]; p->stride[k] = stride[k]; p->padding[k] = padding[k];
	}
	p->pool_type = (char_pool_type != NULL && strcmp(char_pool_type, "AVG") == 0) ? AVG_pool : MAX_pool;
	p->global = global;
	if (previous != NULL && previous->type == POOL) { printf("ERROR: Bad network design, no use of two successive pooling layer.\n"); exit(EXIT_FAILURE); }
	if (previous != NULL && previous->type == DENSE) { printf("ERROR: Unsuported layer types stacking."); exit(EXIT_FAILURE); }
	prev_shape(net, previous, &pc, &ph, &pw);
	pd = prev_depth(net, previous);
	ph /= pd;      /* true height */
	p->prev_size[0] = pw; p->prev_size[1] = ph; p->prev_size[2] = pd; p->prev_depth = pc;
	if (global)
		for (k = 0; k < 3; k++) { p->p_size[k] = p->prev_size[k]; p->stride[k] = p->prev_size[k]; p->padding[k] = 0; }
	if (p->p_size[0] * p->p_size[1] * p->p_size[2] >= 255 && !(global && p->pool_type == AVG_pool && pd == 1)) {
		printf("\nERROR: pooling windows of 255 elements or more are only available as global average pooling of a 2-D map.\n"); exit(EXIT_FAILURE);
	}
	for (k = 0; k < 3; k++)
		p->nb_area[k] = nb_area_comp(p->prev_size[k], p->p_size[k], p->padding[k], 0, p->stride[k]);
	p->nb_maps = pc;
	current->out_c = pc; current->out_w = p->nb_area[0]; current->out_d = p->nb_area[2];
	current->out_h = p->nb_area[1] * p->nb_area[2];
	current->param = p;
	set_activ_defaults(current, activation);
	if (current->activation_type == YOLO) { printf("\nERROR: YOLO activation on a pool layer is not supported.\n"); exit(EXIT_FAILURE); }

	p->desc.dtype = net->dtype; p->desc.batch = net->batch_size; p->desc.length = net->batch_size;
	p->desc.c = pc; p->desc.in_h = ph; p->desc.in_w = pw; p->desc.out_h = p->nb_area[1]; p->desc.out_w = current->out_w;
	if (pd > 1 || p->nb_area[2] > 1 || p->p_size[2] > 1 || p->padding[2] > 0) {      /* 3-D windows: generic kernels (pool.cu) */
		p->desc.in_d = pd; p->desc.out_d = p->nb_area[2]; p->desc.p_d = p->p_size[2]; p->desc.stride_d = p->stride[2]; p->desc.pad_d = p->padding[2];
	}
	p->desc.p_h = p->p_size[1]; p->desc.p_w = p->p_size[0];
	p->desc.stride_h = p->stride[1]; p->desc.stride_w = p->stride[0];
	p->desc.pad_h = p->padding[1]; p->desc.pad_w = p->padding[0];
	p->desc.pool_type = p->pool_type == AVG_pool ? CB200_POOL_AVG : CB200_POOL_MAX;
	p->desc.activ = current->activ;
	if (current->activation_type == SOFTMAX) p->desc.activ.type = CB200_LINEAR;
	if (drop_rate > 0.01f) p->desc.activ.type = CB200_LINEAR;   /* the activation runs in the dropout pass, after the mask */
	cb_dropout_setup(current);

	current->output = dev_alloc(act_bytes(net, current->out_c, current->out_h, current->out_w));
	if (!net->inference_only) {
		current->delta_o = dev_alloc(act_bytes(net, current->out_c, current->out_h, current->out_w));
		if (p->pool_type == MAX_pool)
			p->pool_map = (uint8_t *)dev_alloc((size_t)net->batch_size * current->out_h * current->out_w * cb200_round_channels(pc));
	}
	if (previous != NULL && previous->type == NORM && fusion_on() && (p->pool_map != NULL || net->inference_only)
	    && cb200_norm_pool_fusable(&((norm_param *)previous->param)->desc, &p->desc)) {
		/* group-norm + 2x2 max-pool run as one pass each way; the normalised full-resolution tensor is never stored */
		norm_param *np = (norm_param *)previous->param;
		np->fused_pool = current;
		p->fused_norm = 1;
		CB_CHECK(cb200_free(previous->output)); previous->output = NULL;
		if (previous->delta_o != NULL) { CB_CHECK(cb200_free(previous->delta_o)); previous->delta_o = NULL; }
	}
	current->forward = forward_pool_layer;
	current->backprop = backward_pool_layer;

	print_string_activ_param(current, activ);
	printf("      Input: %dx%dx%dx%d, Output: %dx%dx%dx%d\n"
	       "      P. size: %dx%dx%d, Stride: %dx%dx%d, padding: %dx%dx%d \n"
	       "      Pool type: %s, Global: %d, Activation: %s, dropout rate: %0.2f\n",
		p->prev_size[0], p->prev_size[1], p->prev_size[2], p->prev_depth, p->nb_area[0], p->nb_area[1], p->nb_area[2], p->nb_maps,
		p->p_size[0], p->p_size[1], p->p_size[2], p->stride[0], p->stride[1], p->stride[2],
		p->padding[0], p->padding[1], p->padding[2], p->pool_type == AVG_pool ? "AVG" : "MAX", p->global, activ, current->dropout_rate);
	return net->nb_layers - 1;
}

void pool_save(FILE *f, layer *current, int f_bin)
{
	pool_param *p = (pool_param *)current->param;
	char layer_type = 'P';
	char ptype[40];
	memset(ptype, 0, sizeof(ptype));
	sprintf(ptype, "%s", p->pool_type == AVG_pool ? "AVG" : "MAX");
	if (f_bin) {
		fwrite(&layer_type, sizeof(char), 1, f);
		fwrite(p->p_size, sizeof(int), 3, f);
		fwrite(p->stride, sizeof(int), 3, f);
		fwrite(p->padding, sizeof(int), 3, f);
		fwrite(&p->global, sizeof(int), 1, f);
		fwrite(&current->dropout_rate, sizeof(float), 1, f);
		fwrite(ptype, sizeof(char), 40, f);
		print_activ_param(f, current, f_bin);
	} else {
		fprintf(f, "P%dx%dx%d.%dx%dx%ds%dx%dx%dp%dg_%fd", p->p_size[0], p->p_size[1], p->p_size[2],
			p->stride[0], p->stride[1], p->stride[2], p->padding[0], p->padding[1], p->padding[2], p->global, current->dropout_rate);
		fprintf(f, "%s ", ptype);
		print_activ_param(f, current, f_bin);
		fprintf(f, "\n\n");
	}
}

void pool_load(network *net, FILE *f, int f_bin)
{
	int p_size[3], stride[3], padding[3], global;
	float dropout_rate;
	char pool_type[40], activ_type[40];
	layer *previous;
	printf("Loading pool layer, L:%d\n", net->nb_layers + 1);
	if (f_bin) {
		fread(p_size, sizeof(int), 3, f);
		fread(stride, sizeof(int), 3, f);
		fread(padding, sizeof(int), 3, f);
		fread(&global, sizeof(int), 1, f);
		fread(&dropout_rate, sizeof(float), 1, f);
		fread(pool_type, sizeof(char), 40, f);
		fread(activ_type, sizeof(char), 40, f);
	} else {
		fscanf(f, "%dx%dx%d.%dx%dx%ds%dx%dx%dp%dg_%fd%s %s\n", &p_size[0], &p_size[1], &p_size[2],
			&stride[0], &stride[1], &stride[2], &padding[0], &padding[1], &padding[2], &global, &dropout_rate, pool_type, activ_type);
	}
	previous = net->nb_layers <= 0 ? NULL : net->net_layers[net->nb_layers - 1];
	pool_create(net, previous, p_size, stride, padding, pool_type, activ_type, global, dropout_rate);
}

/* ------------------------------------------------------------------ group normalisation */
static void forward_norm_layer(layer *current)
{
	network *net = current->c_network;
	norm_param *p = (norm_param *)current->param;
	if (net->length == 0) return;
	p->desc.length = net->length;
	if (p->fused_pool != NULL) return;      /* evaluated by the following pool layer (cb200_norm_pool_forward) */
	CB_CHECK(cb200_norm_forward_ex(&p->desc, current->previous->output, current->output, p->gamma, p->beta, p->mean, p->var, p->workspace,
		p->stats_ready, NULL));
	p->stats_ready = 0;
}

static void backward_norm_layer(layer *current)
{
	network *net = current->c_network;
	norm_param *p = (norm_param *)current->param;
	p->desc.length = net->length;
	{
		/* the delta written for a preceding convolution is also column-summed: that is its bias-column gradient */
		layer *prev = current->previous;
		float *colsum = NULL;
		if (prev->type == CONV && !prev->frozen) {
			conv_param *cp = (conv_param *)prev->param;
			if (cp->bias_grad_from_next) colsum = cp->w.grad_b;
		}
		if (p->fused_pool != NULL) {
			layer *pool = p->fused_pool;
			pool_param *pp = (pool_param *)pool->param;
			pp->desc.length = net->length;
			/* (the pool layer's output - masked in place by its dropout, whose backward has already zeroed the same elements of
			 * its delta - lets the backward reductions skip the input-sized tensor) */
			CB_CHECK(cb200_norm_pool_backward_ex(&p->desc, &pp->desc, prev->output, pool->delta_o, pp->pool_map, prev->delta_o,
				p->gamma, p->mean, p->var, p->d_gamma, p->d_beta, &prev->activ, colsum, p->workspace, pool->output, p->beta, NULL));
		} else
			CB_CHECK(cb200_norm_backward(&p->desc, prev->output, current->delta_o, prev->delta_o,
				p->gamma, p->mean, p->var, p->d_gamma, p->d_beta, &prev->activ, colsum, p->workspace, NULL));
	}
	/* data parallel: the batch sums go into the arena now, to be all-reduced before the optimizer; on one GPU the sum is
	 * folded into the parameter update (network.c: cb200_norm_reduce_update) - one tiny launch less per layer */
	if (!current->frozen && cb200_dp_world() > 1)
		CB_CHECK(cb200_norm_reduce_grads(&p->desc, p->d_gamma, p->d_beta, p->gsum, NULL));
}

int norm_create(network *net, layer *previous, const char *norm_type, const char *activation, int group_size, int set_off, FILE *f_load, int f_bin)
{
	layer *current = new_layer(net, NORM, previous);
	norm_param *p;
	float *host;
	int i;
	char activ[40];

	printf("L:%d - CREATING NORMALIZATION LAYER ...\n", net->nb_layers);
	if (previous == NULL) { printf("\nERROR: Normalization layer is not autorized as first layer.\n"); exit(EXIT_FAILURE); }
	if (strncmp(norm_type, "GN", 2) != 0) { printf("\nERROR: Unrecognized normalization type (only GN is available).\n"); exit(EXIT_FAILURE); }
	if (group_size <= 0) { printf("\nERROR: Group Normalization cannot be set with group size <= 0.\n"); exit(EXIT_FAILURE); }
	if (previous->type == DENSE) { printf("\nERROR: normalization layer is not authorized after dense layers atm.\n"); exit(EXIT_FAILURE); }
	if (previous->type == NORM || previous->type == LRN) { printf("\nERROR: stacking two normalization layers is not allowed.\n"); exit(EXIT_FAILURE); }

	p = (norm_param *)calloc(1, sizeof(norm_param));
	if (previous->type == CONV && !((conv_param *)previous->param)->desc.input_is_patches && !drop_on(previous))
		((conv_param *)previous->param)->bias_grad_from_next = 1;   /* (with dropout the conv masks its delta first) */
	p->group_size = group_size; p->set_off = set_off;
	p->n_dim = previous->out_c; p->dim_offset = previous->out_h * previous->out_w;
	p->nb_group = p->n_dim % group_size == 0 ? p->n_dim / group_size : p->n_dim / group_size + 1;
	current->out_c = previous->out_c; current->out_h = previous->out_h; current->out_w = previous->out_w;
	current->param = p;
	load_activ_param(current, activation);
	if (current->activation_type == SOFTMAX) { printf("\nERROR: softmax activation for normalization layer is not authorized\n"); exit(EXIT_FAILURE); }
	if (current->activation_type == YOLO) { printf("\nERROR: YOLO activation for normalization layer is not authorized\n"); exit(EXIT_FAILURE); }
	if (current->activation_type != LINEAR) { printf("\nERROR: only the LIN activation is supported on normalization layers by the B200 core yet.\n"); exit(EXIT_FAILURE); }
	set_activ_defaults(current, activation);

	p->desc.dtype = net->dtype; p->desc.batch = net->batch_size; p->desc.length = net->batch_size;
	p->desc.c = p->n_dim; p->desc.h = current->out_h; p->desc.w = current->out_w;
	p->desc.group_size = group_size; p->desc.nb_group = p->nb_group; p->desc.set_off = set_off; p->desc.eps = 0.001f;

	p->gamma = (float *)dev_alloc(p->nb_group * sizeof(float));
	p->beta = (float *)dev_alloc(p->nb_group * sizeof(float));
	p->mean = (float *)dev_alloc((size_t)p->nb_group * net->batch_size * sizeof(float));
	p->var = (float *)dev_alloc((size_t)p->nb_group * net->batch_size * sizeof(float));
	p->workspace = dev_alloc(cb200_norm_workspace_bytes(&p->desc));
	current->output = dev_alloc(act_bytes(net, current->out_c, current->out_h, current->out_w));
	if (!net->inference_only) {
		p->gamma_update = (float *)dev_alloc(p->nb_group * sizeof(float));
		p->beta_update = (float *)dev_alloc(p->nb_group * sizeof(float));
		p->d_gamma = (float *)dev_alloc((size_t)p->nb_group * net->batch_size * sizeof(float));
		p->d_beta = (float *)dev_alloc((size_t)p->nb_group * net->batch_size * sizeof(float));
		current->delta_o = dev_alloc(act_bytes(net, current->out_c, current->out_h, current->out_w));
	}
	host = (float *)calloc(2 * p->nb_group, sizeof(float));
	for (i = 0; i < p->nb_group; i++) host[i] = 1.0f;
	if (f_load != NULL) {
		if (f_bin) fread(host, sizeof(float), 2 * p->nb_group, f_load);
		else for (i = 0; i < 2 * p->nb_group; i++) fscanf(f_load, "%f", &host[i]);
	}
	CB_CHECK(cb200_h2d(p->gamma, host, p->nb_group * sizeof(float), NULL));
	CB_CHECK(cb200_h2d(p->beta, host + p->nb_group, p->nb_group * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
	free(host);

	current->forward = forward_norm_layer;
	current->backprop = backward_norm_layer;
	current->nb_params = 2 * p->nb_group - set_off;
	print_string_activ_param(current, activ);
	printf("      Group size: %d, Nb. groups: %d, Set-off: %d\n      Activation: %s\n      Nb. params: %d\n",
		p->group_size, p->nb_group, p->set_off, activ, 2 * p->nb_group);
	net->total_nb_param += (2 * p->nb_group - set_off);
	return net->nb_layers - 1;
}

void norm_save(FILE *f, layer *current, int f_bin)
{
	norm_param *p = (norm_param *)current->param;
	char layer_type = 'N';
	char ntype[40];
	float *host = (float *)malloc(2 * p->nb_group * sizeof(float));
	int i;
	memset(ntype, 0, sizeof(ntype));
	sprintf(ntype, "GN");
	if (f_bin) {
		fwrite(&layer_type, sizeof(char), 1, f);
		fwrite(ntype, sizeof(char), 40, f);
		fwrite(&p->group_size, sizeof(int), 1, f);
		fwrite(&p->set_off, sizeof(int), 1, f);
		print_activ_param(f, current, f_bin);
	} else {
		fprintf(f, "N ");
		fprintf(f, "%s ", ntype);
		fprintf(f, "S%d_O%d", p->group_size, p->set_off);
		print_activ_param(f, current, f_bin);
		fprintf(f, "\n");
	}
	CB_CHECK(cb200_d2h(host, p->gamma, p->nb_group * sizeof(float), NULL));
	CB_CHECK(cb200_d2h(host + p->nb_group, p->beta, p->nb_group * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
	if (f_bin) {
		fwrite(host, sizeof(float), 2 * p->nb_group, f);
	} else {
		for (i = 0; i < p->nb_group; i++) fprintf(f, "%g ", host[i]);
		fprintf(f, "\n");
		for (i = 0; i < p->nb_group; i++) fprintf(f, "%g ", host[p->nb_group + i]);
		fprintf(f, "\n\n");
	}
	free(host);
}

void norm_load(network *net, FILE *f, int f_bin)
{
	int group_size, set_off;
	char norm[40], activ_type[40];
	layer *previous;
	printf("Loading norm layer, L:%d\n", net->nb_layers + 1);
	if (f_bin) {
		fread(norm, sizeof(char), 40, f);
		fread(&group_size, sizeof(int), 1, f);
		fread(&set_off, sizeof(int), 1, f);
		fread(activ_type, sizeof(char), 40, f);
	} else {
		fscanf(f, " %s S%d_O%d%s\n", norm, &group_size, &set_off, activ_type);
	}
	previous = net->nb_layers <= 0 ? NULL : net->net_layers[net->nb_layers - 1];
	norm_create(net, previous, norm, activ_type, group_size, set_off, f, f_bin);
}

/* ------------------------------------------------------------------ local response normalisation
 * Same layer as upstream's lrn_create / cuda_forward_lrn_layer / cuda_backward_lrn_layer (src/lrn_layer.c:96-214,
 * src/cuda/cuda_lrn_layer.cu:172-215): no weights, a per-activation scale kept between the two passes. */
static void forward_lrn_layer(layer *current)
{
	network *net = current->c_network;
	lrn_param *p = (lrn_param *)current->param;
	if (net->length == 0) return;
	p->desc.length = net->length;
	CB_CHECK(cb200_lrn_forward(&p->desc, current->previous->output, current->output, p->local_scale, NULL));
}

static void backward_lrn_layer(layer *current)
{
	network *net = current->c_network;
	lrn_param *p = (lrn_param *)current->param;
	layer *prev = current->previous;
	p->desc.length = net->length;
	CB_CHECK(cb200_lrn_backward(&p->desc, prev->output, current->output, current->delta_o, prev->delta_o, p->local_scale,
		&prev->activ, prev->output, NULL));
}

int lrn_create(network *net, layer *previous, const char *activation, int range, float k, float alpha, float beta, FILE *f_load, int f_bin)
{
	layer *current = new_layer(net, LRN, previous);
	lrn_param *p;
	size_t n_act;
	char activ[40];
	(void)f_load; (void)f_bin;

	printf("L:%d - CREATING LOCAL RESPONSE NORMALIZATION LAYER ...\n", net->nb_layers);
	if (previous == NULL) { printf("\nERROR: normalization layer is not autorized as first layer.\n"); exit(EXIT_FAILURE); }
	if (previous->type == DENSE) { printf("\nERROR: normalization layer is not authorized after dense layers atm.\n"); exit(EXIT_FAILURE); }
	if (previous->type == NORM || previous->type == LRN) { printf("\nERROR: stacking two normalization layers is not allowed.\n"); exit(EXIT_FAILURE); }
	if (range <= 0) { printf("\nERROR: LRN range must be > 0.\n"); exit(EXIT_FAILURE); }

	p = (lrn_param *)calloc(1, sizeof(lrn_param));
	p->range = range; p->k = k; p->alpha = alpha; p->beta = beta;
	p->n_dim = previous->out_c; p->dim_offset = previous->out_h * previous->out_w;
	current->out_c = previous->out_c; current->out_h = previous->out_h; current->out_w = previous->out_w;
	current->param = p;
	load_activ_param(current, activation);
	if (current->activation_type == SOFTMAX) { printf("\nERROR: softmax activation for normalization layer is not authorized\n"); exit(EXIT_FAILURE); }
	if (current->activation_type == YOLO) { printf("\nERROR: YOLO activation for normalization layer is not authorized\n"); exit(EXIT_FAILURE); }
	if (current->activation_type != LINEAR) { printf("\nERROR: only the LIN activation is supported on normalization layers by the B200 core yet.\n"); exit(EXIT_FAILURE); }
	set_activ_defaults(current, activation);

	p->desc.dtype = net->dtype; p->desc.batch = net->batch_size; p->desc.length = net->batch_size;
	p->desc.c = p->n_dim; p->desc.h = current->out_h; p->desc.w = current->out_w;
	p->desc.range = range; p->desc.k = k; p->desc.alpha = alpha; p->desc.beta = beta;

	n_act = act_bytes(net, current->out_c, current->out_h, current->out_w);
	current->output = dev_alloc(n_act);
	if (!net->inference_only) {
		p->local_scale = (float *)dev_alloc(n_act / cb200_dtype_size(net->dtype) * sizeof(float));
		current->delta_o = dev_alloc(n_act);
	}
	current->forward = forward_lrn_layer;
	current->backprop = backward_lrn_layer;
	current->nb_params = 0;
	print_string_activ_param(current, activ);
	printf("      Range: %d, k: %f, Alpha: %f, Beta: %f, Activation: %s\n", p->range, p->k, p->alpha, p->beta, activ);
	return net->nb_layers - 1;
}

int cb_lrn_range(layer *current) { return ((lrn_param *)current->param)->range; }

void lrn_save(FILE *f, layer *current, int f_bin)
{
	lrn_param *p = (lrn_param *)current->param;
	char layer_type = 'L';
	if (f_bin) {
		fwrite(&layer_type, sizeof(char), 1, f);
		fwrite(&p->range, sizeof(int), 1, f);
		fwrite(&p->k, sizeof(float), 1, f);
		fwrite(&p->alpha, sizeof(float), 1, f);
		fwrite(&p->beta, sizeof(float), 1, f);
		print_activ_param(f, current, f_bin);
	} else {
		fprintf(f, "L %d %f %f %f ", p->range, p->k, p->alpha, p->beta);
		print_activ_param(f, current, f_bin);
		fprintf(f, "\n");
	}
}

void lrn_load(network *net, FILE *f, int f_bin)
{
	int range;
	float k, alpha, beta;
	char activ_type[40];
	layer *previous;
	printf("Loading Local Response Normalization layer, L:%d\n", net->nb_layers + 1);
	if (f_bin) {
		fread(&range, sizeof(int), 1, f);
		fread(&k, sizeof(float), 1, f);
		fread(&alpha, sizeof(float), 1, f);
		fread(&beta, sizeof(float), 1, f);
		fread(activ_type, sizeof(char), 40, f);
	} else {
		fscanf(f, " %d %f %f %f %s", &range, &k, &alpha, &beta, activ_type);
	}
	previous = net->nb_layers <= 0 ? NULL : net->net_layers[net->nb_layers - 1];
	lrn_create(net, previous, activ_type, range, k, alpha, beta, f, f_bin);
}

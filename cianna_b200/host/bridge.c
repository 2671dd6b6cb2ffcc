/*
 * bridge.c - plain-argument entry points for the upstream-side back-end shim (see cianna_bridge.h).
 *
 * Part of libcianna_host.so (linked -Bsymbolic: the calls below bind to THIS library's init_network / conv_create /
 * ... even when the process also holds upstream's functions of the same names).
 */
#include <math.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>
#include "cianna.h"
#include "cianna_bridge.h"

extern void cb_use_device_batch(network *net, const void *input_dev);
extern void cb_prepare_training(network *net);
extern void cb_set_hyper(network *net, float lr, float momentum, float weight_decay);
extern void cb_apply_updates(network *net);
extern void cb_output_deriv_error(network *net, const void *target_dev);
extern void cb_output_error(network *net, const void *target_dev);
extern void dense_refresh_operands(layer *cur);
extern void dense_get_weights(layer *cur, float *dst, int moment);
extern void dense_set_weights(layer *cur, const float *src);

/* the mirrored layers print their own creation banners; upstream prints its own for the same layer */
static int quiet_fd = -1;
static void quiet_begin(void)
{
	int nul;
	if (getenv("CB200_SHIM_VERBOSE") != NULL) return;
	fflush(stdout);
	quiet_fd = dup(1);
	nul = open("/dev/null", O_WRONLY);
	if (nul >= 0) { dup2(nul, 1); close(nul); }
}
static void quiet_end(void)
{
	if (quiet_fd < 0) return;
	fflush(stdout);
	dup2(quiet_fd, 1);
	close(quiet_fd);
	quiet_fd = -1;
}

static network *net_of(int id)
{
	if (id < 0 || id >= MAX_NETWORKS_NB || networks[id] == NULL) {
		printf("\nERROR: cianna_b200 bridge: network %d does not exist\n", id);
		exit(EXIT_FAILURE);
	}
	return networks[id];
}

static layer *layer_of(network *net, int l)
{
	if (l < 0 || l >= net->nb_layers) { printf("\nERROR: cianna_b200 bridge: layer %d out of range\n", l); exit(EXIT_FAILURE); }
	return net->net_layers[l];
}

void cbb_init(int net_id, const int in_dims[4], int out_dim, float in_bias, int batch_size, int dynamic_load, int tc_mode,
	int inference_only, int adv_size)
{
	static const char *modes[5] = { "FP32C_FP32A", "TF32C_FP32A", "FP16C_FP32A", "FP16C_FP16A", "BF16C_FP32A" };
	int dims[4] = { in_dims[0], in_dims[1], in_dims[2], in_dims[3] };
	unsigned keep = (unsigned)rand();      /* init_network re-seeds rand(); upstream's initialisers must keep their stream */
	quiet_begin();
	init_network(net_id, dims, out_dim, in_bias, batch_size, "C_CUDA", dynamic_load, modes[tc_mode < 0 || tc_mode > 4 ? 0 : tc_mode],
		inference_only, 1, adv_size);
	quiet_end();
	srand(keep);
}

int cbb_nb_layers(int net_id) { return net_of(net_id)->nb_layers; }
int cbb_dtype_size(int net_id) { return (int)cb200_dtype_size(net_of(net_id)->dtype); }

static layer *prev_of(network *net, int prev) { return prev < 0 ? NULL : layer_of(net, prev); }

int cbb_conv(int net_id, int prev, const int f_size[3], int nb_filters, const int stride[3], const int padding[3],
	const int int_padding[3], const char *activation, float bias, float drop_rate, const float *weights, int row_stride)
{
	network *net = net_of(net_id);
	int f[3], s[3], p[3], ip[3], k, l;
	for (k = 0; k < 3; k++) { f[k] = f_size[k]; s[k] = stride[k]; p[k] = padding[k]; ip[k] = int_padding[k]; }
	quiet_begin();
	l = conv_create(net, prev_of(net, prev), f, nb_filters, s, p, ip, NULL, activation, &bias, drop_rate, "xavier", 1.0f, NULL, 0);
	quiet_end();
	if (weights != NULL) {
		conv_param *cp = (conv_param *)net->net_layers[l]->param;
		float *rows = (float *)malloc((size_t)nb_filters * cp->flat_f_size * sizeof(float));
		for (k = 0; k < nb_filters; k++)
			memcpy(rows + (size_t)k * cp->flat_f_size, weights + (size_t)k * row_stride, cp->flat_f_size * sizeof(float));
		cb_layer_set_weights(net, l, rows);
		free(rows);
	}
	return l;
}

int cbb_pool(int net_id, int prev, const int p_size[3], const int stride[3], const int padding[3], int is_avg,
	const char *activation, int global, float drop_rate)
{
	network *net = net_of(net_id);
	int ps[3], s[3], p[3], k, l;
	for (k = 0; k < 3; k++) { ps[k] = p_size[k]; s[k] = stride[k]; p[k] = padding[k]; }
	quiet_begin();
	l = pool_create(net, prev_of(net, prev), ps, s, p, is_avg ? "AVG" : "MAX", activation, global, drop_rate);
	quiet_end();
	return l;
}

int cbb_norm(int net_id, int prev, const char *activation, int group_size, int set_off, const float *gamma, const float *beta)
{
	network *net = net_of(net_id);
	int l;
	quiet_begin();
	l = norm_create(net, prev_of(net, prev), "GN", activation, group_size, set_off, NULL, 0);
	quiet_end();
	if (gamma != NULL && beta != NULL) cbb_norm_set(net_id, l, gamma, beta);
	return l;
}

int cbb_lrn(int net_id, int prev, const char *activation, int range, float k, float alpha, float beta)
{
	network *net = net_of(net_id);
	int l;
	quiet_begin();
	l = lrn_create(net, prev_of(net, prev), activation, range, k, alpha, beta, NULL, 0);
	quiet_end();
	return l;
}

int cbb_dense(int net_id, int prev, int nb_neurons, const char *activation, float bias, float drop_rate, const float *weights)
{
	network *net = net_of(net_id);
	int l;
	quiet_begin();
	/* strict_size = 1: upstream has already applied its own "minus one neuron" alignment rule (src/dense_layer.c:77-84) */
	l = dense_create(net, prev_of(net, prev), nb_neurons, activation, &bias, drop_rate, 1, "xavier", 1.0f, NULL, 0);
	quiet_end();
	if (weights != NULL) cb_layer_set_weights(net, l, weights);
	return l;
}

void cbb_set_yolo(int net_id, int nb_box, int nb_class, int nb_param, int max_nb_obj_per_image, int IoU_type,
	int prior_dist_type, const float *prior_size, const float *noobj_prob_prior, int fit_dim, int strict_box_size,
	int rand_startup, float rand_prob_best_box_assoc, float rand_prob, float min_prior_forced_scaling,
	const float *scale_tab6, const float *slopes_and_maxes_6x3, const float *param_ind_scale, const float *IoU_limits8,
	const int *fit_parts6, int class_softmax, int diff_flag, int error_type, int no_override, int raw_output)
{
	network *net = net_of(net_id);
	yolo_param *y = net->y_param;
	int i;
	if (nb_box <= 0 || nb_box > CB200_YOLO_MAX_BOX) {
		printf("\n ERROR: the B200 core handles 1 to %d YOLO boxes per grid cell (got %d).\n", CB200_YOLO_MAX_BOX, nb_box);
		exit(EXIT_FAILURE);
	}
	/* upstream has resolved defaults / "unset" markers already (src/activ_functions.c:1129-1477): take the values as they are */
	y->no_override = no_override; y->raw_output = raw_output;
	y->nb_box = nb_box; y->nb_class = nb_class; y->nb_param = nb_param; y->max_nb_obj_per_image = max_nb_obj_per_image;
	y->fit_dim = fit_dim; y->IoU_type = IoU_type; y->prior_dist_type = prior_dist_type;
	y->class_softmax = class_softmax; y->diff_flag = diff_flag; y->error_type = error_type;
	y->strict_box_size_association = strict_box_size; y->rand_startup = rand_startup;
	y->rand_prob_best_box_assoc = rand_prob_best_box_assoc; y->rand_prob = rand_prob;
	y->min_prior_forced_scaling = min_prior_forced_scaling;
	y->prior_size = (float *)calloc(3 * (size_t)nb_box, sizeof(float));
	memcpy(y->prior_size, prior_size, 3 * (size_t)nb_box * sizeof(float));
	y->noobj_prob_prior = (float *)calloc(nb_box, sizeof(float));
	memcpy(y->noobj_prob_prior, noobj_prob_prior, nb_box * sizeof(float));
	y->param_ind_scale = (float *)calloc(nb_param > 0 ? nb_param : 1, sizeof(float));
	for (i = 0; i < nb_param; i++) y->param_ind_scale[i] = param_ind_scale != NULL ? param_ind_scale[i] : 1.0f;
	memcpy(y->scale_tab, scale_tab6, 6 * sizeof(float));
	memcpy(y->slopes_and_maxes_tab, slopes_and_maxes_6x3, 18 * sizeof(float));
	memcpy(y->IoU_limits, IoU_limits8, 8 * sizeof(float));
	memcpy(y->fit_parts, fit_parts6, 6 * sizeof(int));
}

void cbb_layer_shape(int net_id, int l, int *c_h_w)
{
	layer *cur = layer_of(net_of(net_id), l);
	c_h_w[0] = cur->out_c; c_h_w[1] = cur->out_h; c_h_w[2] = cur->out_w;
}

float *cbb_master(int net_id, int l)
{
	layer *cur = layer_of(net_of(net_id), l);
	if (cur->type == CONV) return ((conv_param *)cur->param)->w.master;
	if (cur->type == DENSE) return ((dense_param *)cur->param)->w.master;
	return NULL;
}

float *cbb_moment(int net_id, int l)
{
	layer *cur = layer_of(net_of(net_id), l);
	if (cur->type == CONV) return ((conv_param *)cur->param)->w.moment;
	if (cur->type == DENSE) return ((dense_param *)cur->param)->w.moment;
	return NULL;
}

float *cbb_norm_table(int net_id, int l, int what)
{
	layer *cur = layer_of(net_of(net_id), l);
	norm_param *p = (norm_param *)cur->param;
	if (cur->type != NORM) return NULL;
	switch (what) {
	case 0: return p->mean;
	case 1: return p->var;
	case 2: return p->d_gamma;
	case 3: return p->d_beta;
	case 4: return p->gamma;
	case 5: return p->beta;
	default: return NULL;
	}
}

float *cbb_yolo_monitor(int net_id)
{
	network *net = net_of(net_id);
	layer *last = net->net_layers[net->nb_layers - 1];
	if (last->activation_type != YOLO || last->activ_param == NULL) return NULL;
	return ((yolo_param *)last->activ_param)->monitor_dev;
}

void cbb_weights_changed(int net_id, int l)
{
	network *net = net_of(net_id);
	int k;
	for (k = 0; k < net->nb_layers; k++) {
		layer *cur = net->net_layers[k];
		if (l >= 0 && k != l) continue;
		if (cur->type == CONV) {
			conv_param *p = (conv_param *)cur->param;
			CB_CHECK(cb200_conv_prepare_weights(&p->desc, &p->w, NULL));
		} else if (cur->type == DENSE) {
			/* through dense_set_weights: a dense layer above reads this one's pivot weight as a constant */
			dense_param *p = (dense_param *)cur->param;
			float *host = (float *)malloc((size_t)p->in_size * (p->nb_neurons + 1) * sizeof(float));
			dense_get_weights(cur, host, 0);
			dense_set_weights(cur, host);
			free(host);
		}
	}
}

void *cbb_act_ptr(int net_id, int l, int want_delta)
{
	layer *cur = layer_of(net_of(net_id), l);
	if (cur->type == NORM && ((norm_param *)cur->param)->fused_pool != NULL) return NULL;   /* never materialised */
	return want_delta ? cur->delta_o : cur->output;
}

void cbb_norm_set(int net_id, int l, const float *gamma, const float *beta)
{
	network *net = net_of(net_id);
	norm_param *p = (norm_param *)layer_of(net, l)->param;
	CB_CHECK(cb200_h2d(p->gamma, gamma, (size_t)p->nb_group * sizeof(float), NULL));
	CB_CHECK(cb200_h2d(p->beta, beta, (size_t)p->nb_group * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
}

void cbb_norm_get(int net_id, int l, float *gamma, float *beta)
{
	network *net = net_of(net_id);
	norm_param *p = (norm_param *)layer_of(net, l)->param;
	CB_CHECK(cb200_d2h(gamma, p->gamma, (size_t)p->nb_group * sizeof(float), NULL));
	CB_CHECK(cb200_d2h(beta, p->beta, (size_t)p->nb_group * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
}

void cbb_norm_get_async(int net_id, int l, float *gamma, float *beta)
{
	network *net = net_of(net_id);
	norm_param *p = (norm_param *)layer_of(net, l)->param;
	CB_CHECK(cb200_d2h(gamma, p->gamma, (size_t)p->nb_group * sizeof(float), NULL));
	CB_CHECK(cb200_d2h(beta, p->beta, (size_t)p->nb_group * sizeof(float), NULL));
}

void cbb_stream_sync(void) { CB_CHECK(cb200_stream_sync(NULL)); }

/* ------------------------------------------------------------------ one mini-batch, layer by layer */
void cbb_forward_layer(int net_id, int l, const void *input_dev, int length, int is_inference, int mc_model)
{
	network *net = net_of(net_id);
	layer *cur = layer_of(net, l);
	net->length = length;
	net->is_inference = is_inference;
	net->inference_drop_mode = mc_model ? MC_MODEL : AVG_MODEL;
	if (cur->previous == NULL) cb_use_device_batch(net, input_dev);
	cur->forward(cur);
}

void cbb_deriv_output_error(int net_id, const void *target_dev, float TC_scale_factor, int iter, int train_size)
{
	network *net = net_of(net_id);
	cb_set_TC_scale_factor(net, TC_scale_factor);
	net->iter = iter;
	net->train.size = train_size;
	cb_prepare_training(net);
	cb_output_deriv_error(net, target_dev);
}

/* upstream's backprop(l) ends with l's own weight update; here the raw gradients of all layers are applied in one
 * optimizer sweep when the backward sweep reaches the first layer (same result: a layer's update only reads its own
 * gradient, and the data gradients of the sweep were all computed from the pre-update weights on both sides) */
void cbb_backprop_layer(int net_id, int l, float lr, float momentum, float weight_decay, int frozen)
{
	network *net = net_of(net_id);
	layer *cur = layer_of(net, l);
	cur->frozen = frozen;
	if (l == net->nb_layers - 1) {
		cb_prepare_training(net);
		cb_set_hyper(net, lr, momentum, weight_decay);
	}
	cur->backprop(cur);
	if (l == 0) cb_apply_updates(net);
}

void cbb_output_error(int net_id, const void *target_dev, float *err_dev, size_t err_elems)
{
	network *net = net_of(net_id);
	layer *last = net->net_layers[net->nb_layers - 1];
	(void)err_elems;
	if (last->activation_type == YOLO) {
		yolo_param *y = (yolo_param *)last->activ_param;
		cb_output_error(net, target_dev);      /* per-image loss, six-part split and IoU monitor on the device */
		CB_CHECK(cb200_yolo_scatter_parts(err_dev, y->parts_dev, net->batch_size, net->length, last->out_h * last->out_w,
			y->nb_class, y->nb_param, NULL));
	} else
		CB_CHECK(cb200_output_error_elems(err_dev, last->output, target_dev, net->dtype, net->batch_size, net->length,
			last->out_c, last->out_h, last->out_w, last->activation_type == SOFTMAX ? 1 : 0, last->type == DENSE, NULL));
}

void cbb_export_act(int net_id, int l, int want_delta, float *dst_host)
{
	network *net = net_of(net_id);
	if (want_delta) cb_layer_export_delta(net, l, dst_host);
	else cb_layer_export_output(net, l, dst_host);
}

void cbb_sync(void) { CB_CHECK(cb200_device_sync()); }

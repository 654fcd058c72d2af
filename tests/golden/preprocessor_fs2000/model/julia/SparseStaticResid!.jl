function SparseStaticResid!(T::Vector{<: Real}, residual::AbstractVector{<: Real}, y::Vector{<: Real}, x::Vector{<: Real}, params::Vector{<: Real})
@inbounds begin
    residual[1] = y[14] - exp(params[3]);
    residual[2] = -params[5]*log(y[1]) - (1 - params[5])*log(params[4]) + log(y[1]);
    residual[3] = params[2]*y[2]*(params[1]*y[7]^(params[1] - 1)*y[9]^(1 - params[1])*exp(-params[1]*(params[3] + log(y[4]))) + (1 - params[7])*exp(-params[3])/y[4])/(y[15]*y[16]*y[1]) - 1/(y[1]*y[3]);
    residual[4] = -y[10]/y[9] + y[5];
    residual[5] = -params[6]*y[2]*y[3]/((1 - params[6])*(1 - y[9])) + y[10]/y[9];
    residual[6] = -y[2]*y[7]^params[1]*(1 - params[1])*exp(-params[1]*params[3])/(y[5]*y[9]^params[1]) + y[6];
    residual[7] = -params[2]*y[7]^params[1]*y[9]^(1 - params[1])*(1 - params[1])*exp(-params[1]*params[3])/(y[10]*y[1]*y[3]) + 1/(y[2]*y[3]);
    residual[8] = y[3] - y[7]*(1 - params[7])*exp(-params[3]) + y[7] - y[7]^params[1]*y[9]^(1 - params[1])*exp(-params[1]*params[3]);
    residual[9] = -y[1] + y[2]*y[3];
    residual[10] = -y[10] + y[1] + y[8] - 1;
    residual[11] = y[4] - 1;
    residual[12] = y[13] - y[7]^params[1]*y[9]^(1 - params[1])*exp(-params[1]*params[3]);
    residual[13] = y[11] - y[14] + 1;
    residual[14] = y[12] + 1 - y[1]/y[14];
    residual[15] = y[15] - y[3];
    residual[16] = y[16] - y[2];
end
    return nothing
end

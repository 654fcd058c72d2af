function SparseStaticResid!(T::Vector{<: Real}, residual::AbstractVector{<: Real}, y::Vector{<: Real}, x::Vector{<: Real}, params::Vector{<: Real})
@inbounds begin
    residual[1] = (y[2] - y[3] - y[4])/params[1];
    residual[2] = -params[7]*(y[1] - y[5]) + y[3] - y[3]/(params[10]/400 + 1);
    residual[3] = -params[2]*y[2] + y[2] - (1 - params[2])*(params[3]*y[3] + params[4]*(y[1] - y[5]));
    residual[4] = -params[5]*y[5] + y[5];
    residual[5] = -params[6]*y[4] + y[4];
    residual[6] = -params[9] - 100*y[4] + y[6];
    residual[7] = -params[8] - 400*y[3] + y[7];
    residual[8] = -params[10] - params[8] - 4*params[9] - 400*y[2] + y[8];
end
    return nothing
end
